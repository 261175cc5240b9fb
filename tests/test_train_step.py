"""Training step: alphazero_gym_b200.train.Trainer against the unmodified reference's Agent.update.

The goldens (tests/golden/train_*.npz) come from oracle/gen_train_golden.py, which runs ContinuousAgent.update /
DiscreteAgent.update of /root/reference with its own policies, losses and torch optimizers for three consecutive steps.
Bar: every loss component within 1e-5 relative (+1e-6 absolute) at every step; final weights within 2e-6 absolute (the
optimizers divide by sqrt(v) + eps with eps = 1e-10 / 1e-7, so a last-bit difference in a near-zero gradient is a visible
difference in the step: on the GPU, where cuBLAS sums in another order, at most 0.1 % of the weights may deviate more, none by more
than two learning-rate steps, and the update of every other weight must agree in sign and within 5 %; on the CPU no weight may).
"""
import os

import numpy as np
import pytest
import torch

from alphazero_gym_b200.network import PolicyNet
from alphazero_gym_b200.train import LossConfig, Trainer

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "train_a0c_tuned_rmsprop": dict(net=dict(state_dim=3, hidden=128, n_hidden=3, head_dim=6, nonlinearity="elu", num_components=2),
                                    loss=LossConfig(tuned=True, tau=0.1, policy_coeff=0.1, value_coeff=1, alpha=1, reduction="mean"),
                                    opt="rmsprop", clip=0.0),
    "train_a0c_adam_clip": dict(net=dict(state_dim=3, hidden=128, n_hidden=3, head_dim=6, nonlinearity="elu", num_components=2),
                                loss=LossConfig(tuned=False, tau=0.1, policy_coeff=1, value_coeff=1, alpha=1, reduction="mean"),
                                opt="adam", clip=0.5),
    "train_a0c_k1_sum": dict(net=dict(state_dim=3, hidden=128, n_hidden=3, head_dim=2, nonlinearity="elu", num_components=1),
                             loss=LossConfig(tuned=False, tau=0.5, policy_coeff=0.3, value_coeff=2, alpha=0.2, reduction="sum"),
                             opt="adam", clip=0.0),
    "train_a0c_discrete_rmsprop": dict(net=dict(state_dim=4, hidden=128, n_hidden=2, head_dim=2, nonlinearity="relu", num_actions=2),
                                       loss=LossConfig(tuned=False, tau=0.1, policy_coeff=1, value_coeff=1, alpha=1, reduction="mean"),
                                       opt="rmsprop", clip=0.0),
    "train_a0c_discrete_adam": dict(net=dict(state_dim=4, hidden=128, n_hidden=2, head_dim=2, nonlinearity="relu", num_actions=2),
                                    loss=LossConfig(tuned=False, tau=0.1, policy_coeff=1, value_coeff=1, alpha=0.05, reduction="mean"),
                                    opt="adam", clip=1.0),
}


def _run(name, device, cuda_graph=False):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    c = CASES[name]
    torch.manual_seed(0)
    net = PolicyNet(**c["net"]).load_flat(g["w0"]).to(device)
    tr = Trainer(net, c["loss"], optimizer=c["opt"], grad_clip=c["clip"], cuda_graph=cuda_graph)
    infos = []
    for s in range(3):
        batch = dict(obs=torch.from_numpy(g[f"states_{s}"]).to(device), actions=torch.from_numpy(g[f"actions_{s}"]).to(device),
                     counts=torch.from_numpy(g[f"counts_{s}"]).to(device), V_target=torch.from_numpy(g[f"V_{s}"]).to(device))
        infos.append({k: float(v) for k, v in tr.update(batch).items()})
    return g, tr, infos


def _check(name, device, loss_rtol, w_atol, cuda_graph=False, outliers=0.0):
    """outliers: fraction of the weights allowed beyond w_atol.  0 on the CPU (same kernels, same summation order as the reference:
    measured max deviation 2e-7).  On the GPU cuBLAS sums the GEMMs in another order, and RMSprop / Adam divide by sqrt(v) + eps with
    eps = 1e-10 / 1e-7: a last-bit difference in a near-zero gradient is a full-size step, so up to 0.1 % of the weights may differ,
    each by at most two learning-rate steps -- and the UPDATE of every other weight must agree with the reference's in sign and size."""
    g, tr, infos = _run(name, device, cuda_graph)
    for s, info in enumerate(infos):
        for k, v in info.items():
            ref = float(g[f"info_{k}_{s}"])
            assert abs(v - ref) <= loss_rtol * abs(ref) + 1e-6, (name, s, k, v, ref)
    w, ref, w0 = tr.flat_weights().cpu().numpy(), g["w_3"], g["w0"]
    d = np.abs(w - ref)
    far = d > w_atol
    assert far.mean() <= outliers and d.max() <= (2.5e-3 if outliers else w_atol), (name, float(far.mean()), float(d.max()))
    dw, dw_ref = (w - w0)[~far], (ref - w0)[~far]
    moved = np.abs(dw_ref) > 10 * w_atol
    assert np.all(np.sign(dw[moved]) == np.sign(dw_ref[moved]))
    assert np.all(np.abs(dw[moved] - dw_ref[moved]) <= 0.05 * np.abs(dw_ref[moved]) + w_atol)
    if "log_alpha_3" in g.files:
        assert abs(float(tr.loss.log_alpha.detach()) - float(g["log_alpha_3"])) <= 1e-6


@pytest.mark.parametrize("name", sorted(CASES))
def test_update_matches_reference_cpu(name):
    torch.set_num_threads(1)
    _check(name, "cpu", 1e-5, 2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_update_matches_reference_gpu(name):
    """The same step on cuda:0 (cuBLAS GEMMs sum in a different order: losses within 1e-4, weights within 2e-5)."""
    _check(name, "cuda:0", 1e-4, 2e-5, outliers=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_graph_captured_update_matches_reference_gpu(name):
    """The whole step replayed from a CUDA graph (Trainer(cuda_graph=True)): three replays = the reference's three steps, i.e. the
    warm-up steps of the capture leave no trace in weights or optimizer state."""
    _check(name, "cuda:0", 1e-4, 2e-5, cuda_graph=True, outliers=1e-3)


@pytest.mark.gpu
def test_trained_weights_reach_the_search_engine():
    """Trainer.push_weights -> azg_set_weights: a search after the update uses the new network (root V changes, and equals
    the network's own value of the root observation)."""
    from alphazero_gym_b200._cabi import ACT_ELU, CONTINUOUS
    from alphazero_gym_b200.engine import EngineConfig, SearchEngine
    name = "train_a0c_tuned_rmsprop"
    g, tr, _ = _run(name, "cuda:0")
    eng = SearchEngine(EngineConfig(variant=CONTINUOUS, max_rollouts=25, max_trees=64, num_components=2, state_dim=3, hidden=128,
                                    n_hidden=3, activation=ACT_ELU, c_uct=0.05))
    try:
        x = np.stack([np.cos(0.3), np.sin(0.3), 0.5])[None].astype(np.float32)
        eng.set_weights(g["w0"])
        v0, _ = eng.mlp_forward(x)
        tr.push_weights(eng)
        v1, _ = eng.mlp_forward(x)
        with torch.no_grad():
            vt = float(tr.net.value_head(tr.net.trunk(torch.from_numpy(x).cuda())))
        assert abs(float(v1[0]) - vt) <= 1e-5 * abs(vt) + 1e-6 and abs(float(v1[0]) - float(v0[0])) > 1e-5
    finally:
        eng.close()


@pytest.mark.gpu
def test_selfplay_replay_train_loop_on_device():
    """Ranks 1-3 of SURVEY 8(f) together, nothing leaving the GPU: batched self-play steps fill the device replay buffer, the
    trainer consumes shuffled mini-batches from it (graph-replayed steps), the new weights go back into the engine, and the
    next search runs with them.  Checks the bookkeeping (visit counts add up before and after the weight change, losses finite,
    weights moved, the engine evaluates with exactly the trainer's weights)."""
    from alphazero_gym_b200._cabi import ACT_ELU, CONTINUOUS
    from alphazero_gym_b200.engine import EngineConfig, SearchEngine
    from alphazero_gym_b200.network import init_policy_weights
    from alphazero_gym_b200.selfplay import DeviceReplayBuffer, SelfPlayDriver
    B, N = 512, 25
    eng = SearchEngine(EngineConfig(variant=CONTINUOUS, max_rollouts=N, max_trees=B, num_components=2, state_dim=3, hidden=128,
                                    n_hidden=3, activation=ACT_ELU, c_uct=0.05, eval_q8=True))
    try:
        w0 = init_policy_weights(34, 3, 128, 3, 6)
        net = PolicyNet(3, 128, 3, 6, "elu", num_components=2).load_flat(w0).cuda()
        from alphazero_gym_b200.train import root_children
        assert root_children(1.0, 0.5, N) == 5 and eng.cmax >= 5
        tr = Trainer(net, LossConfig(tuned=True), optimizer="rmsprop", cuda_graph=True, n_root_actions=root_children(1.0, 0.5, N))
        tr.push_weights(eng)
        replay = DeviceReplayBuffer(max_size=4 * B, batch_size=256, obs_dim=3, cmax=eng.cmax, device=eng.device)
        drv = SelfPlayDriver(eng, B, N, max_episode_length=200, seed=34, replay=replay)
        for _ in range(3):
            out = drv.step()
            assert int(out["counts"].sum()) == B * N and int(out["n_children"].min()) == 5 and int(out["n_children"].max()) == 5
        assert replay.size == 3 * B
        replay.reshuffle(torch.Generator(device=eng.device).manual_seed(1))
        losses = [float(tr.update(b)["loss"]) for b in replay]
        assert len(losses) == 6 and all(np.isfinite(losses))
        w1 = tr.flat_weights()
        assert float((w1.cpu() - torch.from_numpy(w0)).abs().max()) > 1e-4
        tr.push_weights(eng)
        x = np.array([[np.cos(1.0), np.sin(1.0), -2.0]], np.float32)
        v_eng, _ = eng.mlp_forward(x)
        with torch.no_grad():
            v_net = float(net.value_head(net.trunk(torch.from_numpy(x).cuda())))
        assert abs(float(v_eng[0]) - v_net) <= 1e-5 * abs(v_net) + 1e-6
        out = drv.step()
        assert int(out["counts"].sum()) == B * N
    finally:
        eng.close()
