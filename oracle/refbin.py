"""The reference's own bundled binaries as checkers -- TEST INFRASTRUCTURE.

``tmc3`` (MPEG G-PCC TMC13, the lossless coder of the stride-8 coordinates, ``coder.py:23-36`` through
``gpcc.py:11-36``) and ``pc_error_d`` (MPEG pc_error 0.13.4, the D1 metric, ``pc_error.py:44-54``) ship with the
reference as prebuilt x86-64 executables.  ``__graft_entry__.build()`` copies them to ``oracle/_ref/`` (git-ignored,
travels to the GPU box); this module runs them with exactly the reference's command lines.  Nothing under
``pcgcv2_b200/`` imports this module.
"""
from __future__ import annotations

import os
import re
import subprocess

import numpy as np

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
TMC3 = os.path.join(REF_DIR, "tmc3")
PC_ERROR = os.path.join(REF_DIR, "pc_error_d")


def available() -> bool:
    return os.access(TMC3, os.X_OK) and os.access(PC_ERROR, os.X_OK)


def write_ply(path, coords):
    """``write_ply_ascii_geo`` (``data_utils.py:36-48``) -- same header, same ``x y z`` integer lines."""
    coords = np.asarray(coords).astype("int")
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nend_header\n"
                % coords.shape[0])
        np.savetxt(f, coords, fmt="%d")


def read_ply(path):
    """``read_ply_ascii_geo`` (``data_utils.py:19-34``): lines whose tokens all parse as floats, first three columns."""
    rows = []
    with open(path) as f:
        for line in f:
            try:
                vals = [float(v) for v in line.split(" ") if v != "\n"]
            except ValueError:
                continue
            if len(vals) >= 3:
                rows.append(vals[:3])
    return np.asarray(rows, dtype=np.float64).reshape(-1, 3).astype("int")


def gpcc_encode_coords(coords3, tmpdir) -> bytes:
    """``CoordinateCoder.encode`` (``coder.py:23-29``) with the flags of ``gpcc.py:11-21`` -> contents of ``_C.bin``."""
    ply, out = os.path.join(tmpdir, "c_in.ply"), os.path.join(tmpdir, "c.bin")
    write_ply(ply, coords3)
    subprocess.run([TMC3, "--mode=0", "--positionQuantizationScale=1", "--trisoupNodeSizeLog2=0",
                    "--neighbourAvailBoundaryLog2=8", "--intra_pred_max_node_size_log2=6", "--inferredDirectCodingMode=0",
                    "--maxNumQtBtBeforeOt=4", "--uncompressedDataPath=" + ply, "--compressedStreamPath=" + out],
                   check=True, stdout=subprocess.DEVNULL)
    with open(out, "rb") as f:
        return f.read()


def gpcc_decode_coords(data: bytes, tmpdir) -> np.ndarray:
    """``CoordinateCoder.decode`` (``coder.py:31-36``, ``gpcc.py:30-36``)."""
    binf, ply = os.path.join(tmpdir, "d.bin"), os.path.join(tmpdir, "c_out.ply")
    with open(binf, "wb") as f:
        f.write(data)
    subprocess.run([TMC3, "--mode=1", "--compressedStreamPath=" + binf, "--reconstructedDataPath=" + ply,
                    "--outputBinaryPly=0"], check=True, stdout=subprocess.DEVNULL)
    return read_ply(ply)


def pc_error_d1(a, b, res: int, tmpdir) -> float:
    """``mseF,PSNR (p2point)`` of ``pc_error_d -a A -b B --hausdorff=1 --resolution=res-1`` (``pc_error.py:44-54``,
    read back at ``coder.py:184``)."""
    fa, fb = os.path.join(tmpdir, "a.ply"), os.path.join(tmpdir, "b.ply")
    write_ply(fa, a)
    write_ply(fb, b)
    out = subprocess.run([PC_ERROR, "-a", fa, "-b", fb, "--hausdorff=1", "--resolution=" + str(res - 1)],
                         check=True, capture_output=True, text=True).stdout
    for line in out.splitlines():
        if "mseF,PSNR (p2point)" in line:
            return float(re.findall(r"[-+0-9.eEinfa]+", line.split(":")[-1])[0])
    raise RuntimeError("pc_error_d printed no 'mseF,PSNR (p2point)' line:\n" + out)


# ---- the UNCHANGED reference python files over the drop-in shims ------------------------------------------------------
REF_PY = ("autoencoder.py", "pcc_model.py", "entropy_model.py", "coder.py", "gpcc.py", "pc_error.py")


def reference_sources_installed() -> bool:
    return available() and all(os.path.isfile(os.path.join(REF_DIR, f)) for f in REF_PY)


def load_reference_coder():
    """import the reference's own ``coder`` module (and through it ``pcc_model``, ``autoencoder``, ``entropy_model``,
    ``gpcc``, ``pc_error``) from ``oracle/_ref`` with ``MinkowskiEngine`` / ``torchac`` / ``data_utils`` resolving to the
    drop-in shims -- exactly what a user of the reference does: put the shim directory first on ``sys.path``."""
    import importlib
    import sys

    import pcgcv2_b200
    pcgcv2_b200.install_shims()                                 # shim dir first: MinkowskiEngine, torchac, data_utils
    if REF_DIR not in sys.path:
        sys.path.append(REF_DIR)
    for m in ("coder", "pcc_model", "autoencoder", "entropy_model", "gpcc", "pc_error"):
        sys.modules.pop(m, None)
    coder = importlib.import_module("coder")
    assert os.path.dirname(os.path.abspath(coder.__file__)) == REF_DIR
    assert coder.ME.__version__.endswith("pcgc.b200"), "the reference did not pick up the shim"
    return coder


def reference_state_dict(sd):
    """state_dict for the reference's ``PCCModel.load_state_dict`` (strict, coder.py:142): the fixtures drop the three
    alias entries ``entropy_bottleneck.{matrix,bias,factor}`` (= the last list entries, entropy_model.py:67-80)."""
    out = dict(sd)
    for alias, name in (("matrix", "_matrices"), ("bias", "_biases"), ("factor", "_factors")):
        out[f"entropy_bottleneck.{alias}"] = sd[f"entropy_bottleneck.{name}.3"]
    return out
