"""Row f4: the batched front end applies the reference's RepresentationNetwork (same parameters) to every
(environment, net) block at once; in eval mode its outputs equal the reference's per-net loop."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

from xroute_env_b200.agent import BatchedRepresentationNetwork
from xroute_env_b200.instances import ispd18_geometry, make_instance

REF = "/root/reference/baseline"


def _obs_batch(geom, insts, routed):
    from oracle.oracle import OracleEnv
    envs = [OracleEnv(geom, i) for i in insts]
    for env, nets in zip(envs, routed):
        for n in nets:
            env.step(n)
    obs = [e.obs()[0] for e in envs]                          # [C_e, Z, Y, X] each
    n_rem = torch.tensor([(o.shape[0] - 2) // 7 for o in obs])
    cmax = max(o.shape[0] for o in obs)
    batch = np.zeros((len(obs), cmax) + obs[0].shape[1:], np.float32)
    for k, o in enumerate(obs):
        batch[k, : o.shape[0]] = o
    return obs, torch.from_numpy(batch), n_rem


def test_shapes_and_masking_without_reference():
    torch.manual_seed(0)
    geom = ispd18_geometry(14, 12, 5)
    insts = [make_instance(geom, 5, 5), make_instance(geom, 3, 6)]
    _, batch, n_rem = _obs_batch(geom, insts, [[2], []])
    net = BatchedRepresentationNetwork().eval()
    with torch.no_grad():
        ob, rep, valid = net(batch, n_rem)
    assert ob.shape == (2, 64) and rep.shape == (2, 4, 64) and valid.tolist() == [[True, True, True, True], [True, True, True, False]]
    assert not rep[1, 3].any() and rep[0, 0].abs().sum() > 0
    # chunking must not change anything
    with torch.no_grad():
        _, rep2, _ = net(batch, n_rem, chunk=2)
    assert torch.allclose(rep, rep2, atol=1e-6, rtol=1e-5)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference not mounted")
@pytest.mark.parametrize("shape", [(14, 12, 5), (70, 9, 9)], ids=["small", "wider-than-standard"])
def test_matches_reference_network_in_eval_mode(shape):
    os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
    sys.path.insert(0, REF)
    import baseline_utils as bu
    torch.manual_seed(1)
    ref = bu.RepresentationNetwork(device="cpu")
    sd = ref.state_dict()
    for k, v in sd.items():                                   # non-trivial BatchNorm statistics
        if k.endswith("running_mean"):
            sd[k] = torch.randn_like(v) * 0.1
        if k.endswith("running_var"):
            sd[k] = torch.rand_like(v) + 0.5
    ref.load_state_dict(sd)
    ref.eval()
    ours = BatchedRepresentationNetwork()
    ours.load_state_dict(sd, strict=True)                    # same names, same shapes
    ours.eval()
    geom = ispd18_geometry(*shape)
    insts = [make_instance(geom, 4, 15), make_instance(geom, 3, 16)]
    obs, batch, n_rem = _obs_batch(geom, insts, [[3], []])
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        want_ob, want_nets = ref.forward(obs)
        got_ob, got_rep, valid = ours(batch, n_rem)
    for e in range(2):
        assert torch.allclose(got_ob[e], want_ob[e], atol=1e-5, rtol=1e-4), e
        order = [int(v) for v in obs[e][1].flatten() if v >= 1]
        assert set(order) == set(want_nets[e]) and int(valid[e].sum()) == len(order)
        for r, net_id in enumerate(order):
            assert torch.allclose(got_rep[e, r], want_nets[e][net_id], atol=1e-5, rtol=1e-4), (e, net_id)
