"""bench.py's JSON contract, exercised on CPU through the reference arm (the GPU arm needs a B200)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                       timeout=600)
    return r


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--config", "tiny", "--steps", "2", "--warmup", "1", "--cpu-batch", "4")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "image_text_pairs_per_sec" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] / 1e3 - 4) < 1e-6          # 4 pairs per step
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_under_torchrun_env_only_rank0_reports():
    r = _run("--impl", "reference", "--config", "tiny", "--steps", "1", "--warmup", "0", "--cpu-batch", "2", "--gpus", "2",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_committed_ncu_traffic_table_belongs_to_the_gemm_engine_in_this_tree():
    """bench.py quotes `roofline.traffic` from profiles/ncu_traffic.json only when that capture was taken on a library
    whose GEMM engine is built from the sources in the tree (or on the very binary / source set).  The committed table
    must pass that gate, otherwise the bench line silently loses its DRAM-traffic evidence: re-capture it
    (tools/r2_final2.sh + tools/ncu_traffic.py) after touching csrc/forde_gemm.cuh, forde_api.cu, sm100_ptx.cuh,
    forde_host.h or include/."""
    import __graft_entry__ as G
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
        tj = json.load(f)
    assert tj.get("gemm_src_sha256_16") == G.source_hash(G.GEMM_SOURCES)[:16], \
        "profiles/ncu_traffic.json was captured on different GEMM-engine sources"
    assert set(tj["kernels"]) >= {"sl_up_fwd[50432x768x3072]", "sl_dz_bwd[50432x768x3072]"}
    for k in tj["kernels"].values():
        assert k["traffic_bytes"] > 0 and k["algorithmic_bytes"] > 0
