"""bench.py's host-side logic on the CPU: the prediction-family workload generator, the cost-balanced slicing of ONE
batch over ranks, and both CPU arms of `--impl reference` on the small preset (the oracle port, and -- where the
reference is staged under oracle/_ref -- the unmodified NecessaryPostTrainingEngine) with the keys the contract names."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_prediction_family_generator():
    import bench
    cfg = dict(bench.PRESETS["synthetic_complex_1m"], N=5000, C=200)
    D = 2 * cfg["dim"]
    b = bench.make_jobs(cfg, D, cfg["C"])
    N, (s, p, o), F = cfg["N"], b["pred"], cfg["family"]
    assert len(b["jobs"]) == cfg["C"] + 1 and len(b["jobs"][0]) == F          # job 0 = the base mimic with every fact
    facts_s = b["facts_s"]
    assert ((facts_s[:, 0] == s) ^ (facts_s[:, 2] == s)).all() and o != s     # every fact mentions s exactly once
    base = {tuple(t) for t in b["jobs"][0].tolist()}
    for j in range(1, len(b["jobs"])):
        fj = b["jobs"][j]
        assert cfg["T"][0] <= len(fj) <= cfg["T"][1]
        assert {tuple(t) for t in fj.tolist()} <= base                        # a candidate REMOVES facts (necessary mode)
        assert ((fj[:, 0] == N) ^ (fj[:, 2] == N)).all()                      # the mimic id replaces s
        want = np.unique(np.concatenate([fj[(fj[:, 0] == N) & (fj[:, 1] == p), 2], [o]]))
        np.testing.assert_array_equal(b["filters"][j], want)                  # to_filter[(mimic, p)] after the removal
    again = bench.make_jobs(cfg, D, 50)                                        # the first candidates do not depend on C
    for j in range(1, 20):
        np.testing.assert_array_equal(again["jobs"][j], b["jobs"][j])


def test_one_batch_is_cut_into_balanced_slices():
    import bench
    from kelpie_b200.parallel import shard_bounds
    cfg = dict(bench.PRESETS["synthetic_complex_1m"], N=5000)
    b = bench.make_jobs(cfg, 2 * cfg["dim"], 4096)
    costs = [len(f) for f in b["jobs"][1:]]
    for world in (1, 2, 4, 8):
        bounds = shard_bounds(costs, world)
        assert bounds[0] == 0 and bounds[-1] == 4096
        per = [sum(costs[bounds[r]:bounds[r + 1]]) for r in range(world)]
        assert max(per) <= 1.01 * sum(costs) / world + 64                     # within one candidate of the mean


@pytest.mark.parametrize("port", [True, False])
def test_reference_arm_prints_the_contract_line(port):
    if not port and not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "src")):
        pytest.skip("reference not staged (python oracle/stage_ref.py)")
    env = dict(os.environ, KP_REFERENCE_BUDGET_S="120")
    if port:
        env["KP_REFERENCE_PORT"] = "1"
    # run bench.py as __main__ inside a wrapper that afterwards reports what the process imported / mapped
    wrapper = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'synthetic_complex_small', '--steps', '3', "
               "'--warmup', '1']; runpy.run_path(%r, run_name='__main__'); "
               "print('PRODUCT_IMPORTED', any(m.startswith('kelpie_b200') for m in sys.modules), "
               "'libkelpie_b200' in open('/proc/self/maps').read())" % os.path.join(ROOT, "bench.py"))
    p = subprocess.run([sys.executable, "-c", wrapper], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.strip().splitlines()
    assert lines[-1] == "PRODUCT_IMPORTED False False"  # no module of kelpie_b200 imported, its .so not mapped
    line = json.loads(lines[-2])
    assert line["impl"] == "reference" and line["steps"] == 3 and line["unit"] == "candidates/s" and line["higher_is_better"]
    assert line["cpu_baseline"]["kind"] == ("port" if port else "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1000.0) < 1e-6
