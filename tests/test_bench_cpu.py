"""CPU: bench.py's bookkeeping -- the algorithmic FLOP figures SURVEY 8d quotes, the workload
selection, and the reference arm's JSON contract (the product arm needs a B200)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_flops_per_image_match_the_survey():
    # cfg2 293.3, cfg3 2286.3, cfg4 (M = 154) 7012.3 GFLOP per image (train = 3 x forward)
    f2, a2 = bench.train_flops_per_image(bench.CFG2, 256, 154)
    assert f2 / 1e9 == pytest.approx(293.3, rel=2e-3)
    assert a2 / f2 == pytest.approx(0.063, abs=2e-3)           # attention share 6.3 %
    big = dict(bench.CFG2, dim=1536, num_heads=24, num_blocks=24)
    assert bench.train_flops_per_image(big, 256, 154)[0] / 1e9 == pytest.approx(2286.3, rel=2e-3)
    assert bench.train_flops_per_image(big, 1024, 154)[0] / 1e9 == pytest.approx(7012.3, rel=2e-3)


def test_config_table_matches_baseline_configs():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert "depth 12 / dim 768" in base["configs"][1] and "batch 64" in base["configs"][1]
    over, latent, batch, label = bench.CONFIGS["cfg2"]
    assert (over, latent, batch) == ({}, 32, 64) and "configs[1]" in label
    assert bench.CONFIGS["cfg3"][0] == dict(dim=1536, num_heads=24, num_blocks=24)
    assert bench.CONFIGS["cfg4"][1] == 64                       # 64x64 latent = 512 px


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "images/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
