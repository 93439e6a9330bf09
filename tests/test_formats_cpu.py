"""On-disk formats: policy checkpoint round trip incl. the fp16 cast path, teacher-pair files."""
import os
import random

import pytest
import torch

import consolver_b200 as cb
from consolver_b200 import formats
from golden_io import Golden


def test_checkpoint_written_by_the_reference_layout_loads_and_round_trips(tmp_path):
    g = Golden("sd_eps_s0_n8_B3")                       # weights that came out of the reference's FactorNetPPO
    ref_path = tmp_path / "checkpoint-10" / "model.ckpt"
    os.makedirs(ref_path.parent)
    torch.save(g.state_dict, ref_path)                  # what train_ppo.py:177 writes
    s = cb.PPOScheduler(order_dim=4, scaler_dim=0, factor_net_kwargs=dict(embedding_dim=64, hidden_dim=256, num_actions=11))
    formats.load_policy(s.factor_net, str(ref_path))
    for k, v in s.factor_net.state_dict().items():
        assert torch.equal(v, g.state_dict[k])
    out = tmp_path / "out" / "model.ckpt"
    formats.save_policy(s.factor_net, str(out))
    again = torch.load(out, map_location="cpu", weights_only=True)
    assert list(again.keys()) == list(g.state_dict.keys())
    assert all(torch.equal(again[k], g.state_dict[k]) for k in again)
    hp = formats.policy_hparams_from_ckpt(str(out))
    assert hp == dict(hidden_dim=256, num_actions=11, action_dims=3, scaler_dim=0, order_dim=4, use_conv=False)
    # gen_ppo.py:193-195: cast the loaded module (and its bin buffer) to fp16
    formats.load_policy(s.factor_net, str(out), dtype=torch.float16)
    assert s.factor_net.mlp[0].weight.dtype == torch.float16 and s.factor_net.action_values.dtype == torch.float16


def test_checkpoint_mismatch_is_reported(tmp_path):
    g = Golden("sd_eps_s2_n8_B3")                       # scaler_dim=2 -> A=5
    p = tmp_path / "model.ckpt"
    torch.save(g.state_dict, p)
    s = cb.PPOScheduler(order_dim=4, scaler_dim=0, factor_net_kwargs=dict(hidden_dim=256, num_actions=11))
    with pytest.raises(ValueError, match="shape mismatch"):
        formats.load_policy(s.factor_net, str(p))
    torch.save({"foo": torch.zeros(1)}, p)
    with pytest.raises(ValueError, match="not a ConsistencySolver"):
        formats.load_policy(s.factor_net, str(p))


def test_teacher_pair_files_round_trip(tmp_path):
    d = str(tmp_path / "pairs")
    noise, latent = torch.randn(4, 64, 64), torch.randn(4, 64, 64)
    formats.save_sd_pair(d, "0_00000007", "a photo of a cat", noise, latent)
    open(os.path.join(d, "0_00000008.txt"), "w").write("incomplete pair")
    assert formats.sd_pair_ids(d) == ["0_00000007"]
    text, n2, l2 = formats.load_sd_pair(d, "0_00000007")
    assert text == "a photo of a cat" and torch.equal(n2, noise) and torch.equal(l2, latent)
    bad = latent.clone()
    bad[0, 0, 0] = float("nan")
    formats.save_sd_pair(d, "0_00000009", "x", noise, bad)
    with pytest.raises(ValueError, match="NaN"):
        formats.load_sd_pair(d, "0_00000009")
    root = tmp_path / "flux"
    os.makedirs(root / "initial_noises"), os.makedirs(root / "obtained_noises")
    a, b = torch.randn(1, 4096, 64).bfloat16(), torch.randn(1, 4096, 64).bfloat16()
    torch.save(a, root / "initial_noises" / "3.pt"), torch.save(b, root / "obtained_noises" / "3.pt")
    a2, b2 = formats.load_flux_pair(str(root), 3)
    assert torch.equal(a, a2) and torch.equal(b, b2)


def test_repeat_random_sample_replicates_one_element():
    noise, target = torch.randn(5, 4, 8, 8), torch.randn(5, 4, 8, 8)
    n, t, texts, i = formats.repeat_random_sample(noise, target, list("abcde"), random.Random(3))
    assert n.shape == noise.shape and all(torch.equal(n[j], noise[i]) for j in range(5))
    assert all(torch.equal(t[j], target[i]) for j in range(5)) and texts == ["abcde"[i]] * 5
