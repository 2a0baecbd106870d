"""-m "not gpu": the CPU legs of bench.py run the REFERENCE'S OWN files (oracle/ref_timing.py).  Here: the copy that travels
to the GPU box (baseline/_ref) is byte-identical to /root/reference where that tree exists, and both drivers play hands and
count transitions as the reference's own counters do."""
import filecmp
import os

import pytest

from oracle import ref_timing


def test_travelling_copy_is_verbatim():
    src = "/root/reference"
    if not os.path.isfile(os.path.join(src, "leduc", "newenv.py")):
        pytest.skip("no reference tree on this machine (the GPU box): the copy under baseline/_ref is all there is")
    assert ref_timing.materialise()
    for rel in ref_timing.FILES:
        assert filecmp.cmp(os.path.join(src, rel), os.path.join(ref_timing.REF_COPY, rel), shallow=False), rel


@pytest.mark.parametrize("kind", ["legacy", "nfsp"])
def test_drivers_play_and_count(kind):
    if ref_timing.reference_root() is None:
        pytest.skip("no reference tree")
    r = ref_timing.run(kind, 300, 1)
    assert r["hands_per_proc"] == 300 and r["seconds"] > 0
    # a Leduc hand is 1-6 decisions under main.train; the README loop steps both players per iteration (2-10 per hand)
    lo, hi = (2, 10) if kind == "legacy" else (1, 6)
    assert lo * 300 <= r["transitions"] <= hi * 300
    assert r["transitions_per_sec"] == pytest.approx(r["transitions"] / r["seconds"])
