"""The arithmetic of the pre-split half-precision kernels (x = f16 hi + f16 lo, three products, lo*lo dropped, weights
split after a power-of-two scale), emulated on the oracle over the whole network (tools/h2_emulation.py): it must stay
two orders of magnitude inside the 1e-4 activation bar and leave bitstream and decoded set unchanged.  CPU only."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_h2_arithmetic_emulated_over_the_network():
    import h2_emulation
    worst, peak, same_stream, same_set = h2_emulation.run("r3", "cube32", verbose=False)
    assert worst < 5e-6, f"h2 emulation: worst per-layer relative error {worst:.2e}"
    assert peak < 6.0e4, "activations leave the f16 range"
    assert same_stream and same_set
