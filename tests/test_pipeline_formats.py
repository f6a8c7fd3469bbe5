"""CPU: the artefact writers of ``pipeline.py`` reproduce the files the reference's own ``main()``s wrote
(tests/golden/pipeline/, made by oracle/make_golden_pipeline.py from 08 / 10 / 9105 unmodified)."""
import csv
import json
import re
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

from amphibian_vae_latent_detector_b200 import pipeline

GOLD = Path(__file__).parent / "golden" / "pipeline"
GRID = ("0.10", "0.15", "0.20", "0.25")
# 9105_make_config_snapshot_from_log.py:11-13
RK_LINE = re.compile(r"✅\s+(?P<sp>[\w_]+):\s+rk_in=(?P<rk_in>[0-9.]+)\s+\|\s+rk_out=(?P<rk_out>[0-9.]+)\s+\|\s+rk=(?P<rk>[0-9.]+)")


def _rows(q):
    rows = []
    with open(GOLD / f"qout_{q}" / "results.csv", newline="", encoding="utf-8") as f:
        for r in csv.DictReader(f):
            rows.append({"file": r["file"], "true_species": r["true_species"], "pred_species": r["pred_species"],
                         "detected": r["detected"] == "True", "correct": r["correct"] == "True",
                         "best_distance": float(r["best_distance"])})
    return rows


@pytest.mark.parametrize("q", GRID)
def test_summary_txt_is_byte_identical(tmp_path, q):
    out = tmp_path / "summary.txt"
    pipeline.write_summary(_rows(q), out)
    assert out.read_text(encoding="utf-8") == (GOLD / f"qout_{q}" / "summary.txt").read_text(encoding="utf-8")


@pytest.mark.parametrize("q", GRID)
def test_results_csv_is_byte_identical(tmp_path, q):
    rows = _rows(q)
    res = [(r["detected"], None if r["pred_species"] == "NO_DETECT" else r["pred_species"], r["best_distance"]) for r in rows]
    again = pipeline._results_rows([r["file"] for r in rows], [r["true_species"] for r in rows], res)
    pipeline._write_results(again, tmp_path)
    assert (tmp_path / "results.csv").read_text(encoding="utf-8") == (GOLD / f"qout_{q}" / "results.csv").read_text(encoding="utf-8")


def _fake_fit(cfgs):
    """A RadialFit-shaped object holding the reference's numbers for the whole grid."""
    names = list(cfgs[0]["radial_detector"]["centroids"].keys())
    per = [c["radial_detector"]["meta_fit"]["per_species"] for c in cfgs]
    summ = lambda d: np.array([d["min"], d["p50"], d["p90"], d["max"]])
    return names, SimpleNamespace(
        centroids=np.array([cfgs[0]["radial_detector"]["centroids"][sp] for sp in names], dtype=np.float32),
        rk_in=np.array([per[0][sp]["rk_in"] for sp in names]),
        rk_out=np.array([[p[sp]["rk_out"] for sp in names] for p in per]),
        rk=np.array([[p[sp]["rk_final"] for sp in names] for p in per]),
        summaries={"in": [summ(per[0][sp]["rho_in_summary"]) for sp in names],
                   "out": [summ(per[0][sp]["rho_out_summary"]) for sp in names]})


def test_config_block_and_log_lines_match_reference():
    cfgs = [json.loads((GOLD / f"qout_{q}" / "config_used.json").read_text(encoding="utf-8")) for q in GRID]
    names, fit = _fake_fit(cfgs)
    for qi, q in enumerate(GRID):
        ref = cfgs[qi]["radial_detector"]
        mf = ref["meta_fit"]
        Z_by = {sp: np.zeros((mf["per_species"][sp]["N_in"], 4), np.float32) for sp in names}
        lines = []
        block = pipeline.radial_config_block(
            names, fit, qi, Z_by, {sp: 0 for sp in names}, {sp: mf["per_species"][sp]["used"] for sp in names},
            Path(mf["chunks_dir"]), mf["chunk_seconds"], q_in=mf["q_in"], q_out=mf["q_out"],
            max_per_class=mf["max_per_class"], seed=mf["seed"], mel=pipeline.MEL_DEFAULTS, log=lines.append)
        assert block == ref                                   # same keys, order-insensitive, same numbers
        assert list(block["meta_fit"].keys()) == list(mf.keys())
        assert list(block["meta_fit"]["per_species"][names[0]].keys()) == list(mf["per_species"][names[0]].keys())
        ref_log = (GOLD / f"qout_{q}" / "run.log").read_text(encoding="utf-8").splitlines()
        want = [ln for ln in ref_log if RK_LINE.search(ln) or ln.startswith("   rho_")]
        assert lines == want
        snap = json.loads((GOLD / f"qout_{q}" / "config_snapshot.json").read_text())
        for ln in lines:
            m = RK_LINE.search(ln)
            if m:
                assert float(m.group("rk")) == snap["rk_per_species"][m.group("sp")]
