"""CPU: host-side logic -- synthetic inputs, sharding, histogram reduction over gloo (world_size 2), weight packing."""
import os
import subprocess
import sys

import numpy as np
import torch

from excel_b200 import evaluate, synth
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_is_deterministic_and_shaped():
    a, b = synth.images(2, 64, seed=5), synth.images(2, 64, seed=5)
    assert torch.equal(a, b) and a.shape == (2, 3, 64, 64)
    lab = synth.class_labels(64, 20, seed=1, n_fixed=None)
    n = lab.sum(1)
    assert n.min() >= 1 and n.max() <= 6 and 1.2 < n.mean() < 2.0      # VOC: mean 1.55, max 6
    t = synth.text_bank(45, 512, seed=0)
    assert torch.allclose(t.norm(dim=1), torch.ones(45), atol=1e-6)


def test_shard_indices_cover_exactly_once():
    for n, w in ((10, 1), (10, 2), (17, 8), (3, 4)):
        got = sorted(i for r in range(w) for i in evaluate.shard_indices(n, r, w))
        assert got == list(range(n))


def test_scores_match_oracle_hist():
    rng = np.random.default_rng(0)
    lt, lp = rng.integers(0, 21, (4, 50, 50)), rng.integers(0, 21, (4, 50, 50))
    lt[0, :5] = 255
    hist = port.fast_hist(lt, lp, 21)
    s = evaluate.scores_from_hist(torch.from_numpy(hist))
    assert abs(s["miou"] - port.miou_from_hist(hist)) < 1e-12


def test_pack_from_visual_roundtrip():
    """encoder.pack_from_visual understands a CLIP-style module tree before and after the surgery swap."""
    from excel_b200.encoder import pack_from_visual

    class Blk(torch.nn.Module):
        def __init__(self, d, surgery):
            super().__init__()
            if surgery:
                self.attn = torch.nn.Module()
                self.attn.qkv, self.attn.proj = torch.nn.Linear(d, 3 * d), torch.nn.Linear(d, d)
            else:
                self.attn = torch.nn.MultiheadAttention(d, d // 64)
            self.ln_1, self.ln_2 = torch.nn.LayerNorm(d), torch.nn.LayerNorm(d)
            self.mlp = torch.nn.Sequential()
            self.mlp.add_module("c_fc", torch.nn.Linear(d, 4 * d))
            self.mlp.add_module("c_proj", torch.nn.Linear(4 * d, d))

    class Vis(torch.nn.Module):
        def __init__(self, d=64, L=3):
            super().__init__()
            self.conv1 = torch.nn.Conv2d(3, d, 16, 16, bias=False)
            self.class_embedding = torch.nn.Parameter(torch.randn(d))
            self.positional_embedding = torch.nn.Parameter(torch.randn(5, d))
            self.ln_pre, self.ln_post = torch.nn.LayerNorm(d), torch.nn.LayerNorm(d)
            self.proj = torch.nn.Parameter(torch.randn(d, 32))
            self.transformer = torch.nn.Module()
            self.transformer.resblocks = torch.nn.Sequential(*[Blk(d, i == L - 1) for i in range(L)])
            self.num_heads = 1

    W = pack_from_visual(Vis())
    assert [int(v) for v in W["meta"]] == [3, 1, 16]
    assert W["blocks.2.in_proj_weight"].shape == (192, 64) and W["blocks.0.out_proj.bias"].shape == (64,)
    assert set(k.split(".", 2)[2] for k in W if k.startswith("blocks.0.")) == set(k.split(".", 2)[2] for k in W if k.startswith("blocks.2."))


def test_hist_all_reduce_gloo_world2(tmp_path):
    """The path's only collective, on CPU with gloo and 2 ranks: sharded histograms sum to the single-rank one."""
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {ROOT!r})
from excel_b200 import evaluate
from oracle import port
dist.init_process_group('gloo')
r, w = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(7)
lt, lp = rng.integers(0, 21, (9, 40, 40)), rng.integers(0, 21, (9, 40, 40))
mine = evaluate.shard_indices(9, r, w)
hist = torch.from_numpy(sum(port.fast_hist(lt[i], lp[i], 21) for i in mine)).to(torch.int64)
evaluate.all_reduce_hist(hist)
full = sum(port.fast_hist(lt[i], lp[i], 21) for i in range(9))
assert np.array_equal(hist.numpy(), full), r
dist.destroy_process_group()
print('ok', r)
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_install_rebinds_reference_symbols():
    """Drop-in mechanics (needs the reference tree: build container only)."""
    import pytest
    from oracle import ref_harness as H
    if not H.available():
        pytest.skip("reference tree not present (GPU box)")
    H.load()
    import importlib
    from excel_b200 import install as inst, par, affutils
    orig = inst.install()
    try:
        aff = importlib.import_module("utils.affutils")
        assert importlib.import_module("utils.PAR").PAR is par.PAR
        assert aff.refine_cams_with_bkg_weclip is affutils.refine_cams_with_bkg_weclip
        ns = {}
        exec("from utils.affutils import refine_cams_with_aff, refine_cams_with_bkg_weclip\nfrom utils.PAR import PAR\nimport clip", ns)
        assert ns["PAR"] is par.PAR and ns["clip"].clip_feature_surgery.__module__ == "excel_b200.clip"
        # a CPU tensor must not silently run the reference: the patched path refuses (no fallback)
        with pytest.raises(RuntimeError):
            ns["clip"].clip_feature_surgery(torch.zeros(1, 5, 8), torch.zeros(3, 8))
    finally:
        inst.uninstall(orig)
    assert importlib.import_module("utils.PAR").PAR is not par.PAR


def test_merge_flipped_maps_has_no_cpu_path():
    import pytest
    from excel_b200.camutils import merge_flipped_maps
    with pytest.raises(RuntimeError):
        merge_flipped_maps(torch.rand(4, 36, 20), 2, 6, 6)


def test_segments_group_runs_of_equal_plane_count():
    from excel_b200.affutils import _segments
    assert _segments([2, 2, 3, 3, 3, 4, 5, 7]) == [(0, 2, 2), (2, 5, 3), (5, 6, 4), (6, 8, 7)]
    assert _segments([3]) == [(0, 1, 3)]
    assert _segments([6, 6, 9]) == [(0, 3, 9)]       # everything above 4 planes shares one run (passes of 3 planes)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path through the oracle port): one JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0 and line["higher_is_better"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("ViT-B/16 CAM+SVC+PAR")


def test_bench_cuda_arm_fails_loudly_without_gpu():
    """No GPU here: the product arm must raise, not fall back to a CPU path."""
    import subprocess
    import sys
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "images/s" not in r.stdout
