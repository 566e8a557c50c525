"""Host logic of FramePipeline (frames in flight: worker threads, job routing, error propagation) with the codec and
the CUDA stream objects replaced by stand-ins -- no GPU needed; the GPU run of the same class is
tests/test_codec_gpu.py::test_frame_pipeline_matches_single_frame_path."""
import threading
import time

import numpy as np
import pytest
import torch

import pcgcv2_b200.pipeline as P


class _FakeStream:
    def wait_stream(self, other):
        pass


class _Ctx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _FakeCodec:
    instances = []

    def __init__(self, *a, **k):
        self.frames, self.threads = [], set()
        _FakeCodec.instances.append(self)

    def encode(self, frame):
        if frame is None:
            raise ValueError("bad frame")
        self.frames.append(int(frame[0]))
        self.threads.add(threading.get_ident())
        time.sleep(0.002)                                   # the other worker must be able to run meanwhile
        return ("stream", int(frame[0]))

    def decode(self, st, rho=1.0, to_host=True):
        return np.full(3, st[1] * (2 if rho == 2.0 else 1))


@pytest.fixture
def fake_cuda(monkeypatch):
    _FakeCodec.instances = []
    monkeypatch.setattr(P, "Codec", _FakeCodec)
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: _FakeStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: _Ctx())
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _FakeStream())
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)


def test_roundtrip_order_routing_and_errors(fake_cuda):
    with P.FramePipeline({}, "cuda", depth=2) as pipe:
        assert pipe.device == torch.device("cuda", 0)       # an index-less device is resolved once, for the workers
        frames = [np.array([i]) for i in range(7)]
        res = pipe.roundtrip(frames, rho=2.0)
        assert [st[1] for st, _ in res] == list(range(7))   # results come back in input order
        assert all((out == 2 * i).all() for i, (_, out) in enumerate(res))
        a, b = _FakeCodec.instances
        assert a.frames == [0, 2, 4, 6] and b.frames == [1, 3, 5]         # frame i -> worker i % depth, in order
        assert len(a.threads) == 1 and len(b.threads) == 1 and a.threads != b.threads
        with pytest.raises(ValueError):                     # a worker's exception reaches the caller ...
            pipe.roundtrip([np.array([1]), None, np.array([3])])
        assert len(pipe.roundtrip([np.array([9])])) == 1    # ... and the pipeline stays usable
        assert pipe.roundtrip([]) == []
    assert not any(t.is_alive() for t in pipe._threads)


def test_bad_arguments(fake_cuda):
    with pytest.raises(ValueError):
        P.FramePipeline({}, "cuda", depth=0)
    with pytest.raises(ValueError):
        P.FramePipeline({}, "cpu")
