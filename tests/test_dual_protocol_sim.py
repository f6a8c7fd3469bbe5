"""CPU: the synchronisation protocol of csrc/dftf4.cu (opt-in dual-tile STFT GEMM, not yet run on hardware) under the
host-side model of tools/dual_protocol_sim.py -- randomised, adversarially paced schedules must finish without deadlock,
ring or TMEM hazard; deliberately broken variants of the protocol must be caught (so a pass means something)."""
import importlib.util
from pathlib import Path

import pytest

SIM = Path(__file__).resolve().parents[1] / "tools" / "dual_protocol_sim.py"
SRC = SIM.read_text()


def _load(src=SRC):
    ns = {"__name__": "dual_protocol_sim_under_test"}
    exec(compile(src, str(SIM), "exec"), ns)
    return ns


@pytest.mark.parametrize("npairs", [1, 2])
def test_protocol_survives_random_schedules(npairs):
    """npairs = 1: AVLD_DFT_DUAL=1; npairs = 2: the variant that shares B by TMA multicast inside a 4-CTA cluster."""
    ns = _load()
    for seed in range(8):
        assert ns["Sim"](3, seed, npairs).run()


def test_model_constants_follow_the_kernel():
    cu = (SIM.parents[1] / "amphibian_vae_latent_detector_b200" / "csrc" / "dftf4.cu").read_text()
    ns = _load()
    assert "mbar_init(&b_empty[s], kPairs)" in cu and "umma_commit_pair(&b_empty[sb], all_mask)" in cu
    assert f"kSA = {ns['K_SA']}, kSB = {ns['K_SB']}" in cu and f"kRegions = {ns['REGIONS']}" in cu
    assert f"kEpiWarps = {ns['EPI_WARPS']}" in cu
    # the region table is the one compiled into the kernel
    assert "if (g == 0) return t == 0 ? (part == 0 ? 0 : 2) : (part == 0 ? 1 : 0);" in cu
    assert "if (g == 1) return part == 0 ? 2 : 1;" in cu and "return part == 0 ? 0 : 2;" in cu
    table = {(g, t, p): ns["region_of"](g, t, p) for g in range(3) for t in range(2 if g == 0 else 1) for p in range(2)}
    assert table == {(0, 0, 0): 0, (0, 0, 1): 2, (0, 1, 0): 1, (0, 1, 1): 0, (1, 0, 0): 2, (1, 0, 1): 1, (2, 0, 0): 0, (2, 0, 1): 2}


@pytest.mark.parametrize("name,old,new", [
    ("issuer skips the drain wait",
     'while not self.r_empty[q][r].passed(((used >> r) & 1) ^ 1, ("issuer", q), "empty"):\n                                    yield', "pass"),
    ("epilogue reads Im1 from the wrong region", "epilogue_region_of = region_of",
     "epilogue_region_of = lambda g, t, part: 1 if (g, t, part) == (0, 1, 1) else region_of(g, t, part)"),
    ("producer skips the B-slot wait",
     'while not self.b_empty[q][cta][sb].passed(pb ^ 1, ("prod", q, cta), "empty"):\n                            yield', "pass"),
    ("r_empty counts one CTA's warps only", 'B(f"r_empty{q}.{r}", 2 * EPI_WARPS)', 'B(f"r_empty{q}.{r}", EPI_WARPS)'),
    ("a B slot is released by one issuer's commit alone", 'B(f"b_empty{q}.{c}.{s}", P)', 'B(f"b_empty{q}.{c}.{s}", 1)'),
    ("B commits reach the own pair only",
     "[self.b_empty[q2][c][sb] for q2 in range(P) for c in range(2)]", "[self.b_empty[q][c][sb] for c in range(2)]"),
    ("epilogue never flips its parity bit", "                        used ^= 1 << r\n                        yield\n                    self.outputs",
     "                        yield\n                    self.outputs"),
])
def test_broken_protocols_are_caught(name, old, new):
    assert old in SRC, name
    ns = _load(SRC.replace(old, new))
    caught = 0
    for seed in range(6):
        try:
            ns["Sim"](3, seed, 2).run()
        except AssertionError:
            caught += 1
    assert caught >= 5, name
