"""The py3 restatements of the reference drivers (src/pose_train.py, src/depth_train.py) run end to end on the device
trainer, and the three phases chain through their snapshots the way the reference's do (pose_train.py:183-184 save_vae ->
depth_train.py:118-129 load_vae / resume)."""
import glob
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "src", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_three_phase_pipeline_chains_through_snapshots(tmp_path, capsys):
    prefix = str(tmp_path / "out" / "pre")
    cfg = os.path.join(ROOT, "exps", "nnyu.yaml")
    pose = _load("pose_train").main(["pose_train.py", "--config", cfg, "--iters", "3", "--batch", "4", "--snapshot_prefix", prefix])
    assert float(pose.vae_total_loss) > 0
    vae_files = glob.glob(prefix + "_vae_3.00_*.pkl")                 # saved as 2 + frac (pose_train.py:184)
    assert len(vae_files) == 1, vae_files
    dt = _load("depth_train")
    pre = dt.main(["depth_train.py", "--config", cfg, "--mode", "pretrain", "--iters", "3", "--batch", "2", "--noise", "host",
                   "--snapshot_prefix", prefix])
    assert all(float(getattr(pre, k)) == float(getattr(pre, k)) for k in ("dis_loss", "gen_total_loss"))   # finite
    assert glob.glob(prefix + "_gen_00000003.pkl") and glob.glob(prefix + "_dis_00000003.pkl")
    gen_after_pretrain = {k: v.clone() for k, v in pre.gen_store.state_dict().items()}
    capsys.readouterr()
    est = dt.main(["depth_train.py", "--config", cfg, "--mode", "estimate3", "--iters", "3", "--noise", "host",
                   "--snapshot_prefix", prefix, "--augment", "1", "--eval_every", "3"])     # GPU augmentation + device eval sweep
    out = capsys.readouterr().out
    assert "Failed to load the parameters of vae" not in out       # estimate3 loads vae_%.2f with 2 + frac (depth_train.py:120)
    # estimate phases start from the pretrained networks (depth_train.py:126-128) and never step the generator
    for k, v in est.gen_store.state_dict().items():
        assert torch.equal(v, gen_after_pretrain[k]), k
    for k, v in est.vae_store.state_dict().items():
        assert torch.equal(v, pose.vae_store.state_dict()[k]), k
    assert float(est.dis_total_loss) > 0 and glob.glob(prefix + "_est_dis_00000003.pkl")
    mean_err, within = est.last_eval                                # the every-N-iterations test sweep ran on the device
    assert mean_err == mean_err and 0.0 <= within <= 100.0 and "Mean err" in out
