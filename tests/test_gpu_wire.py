"""Row f2 on the GPU: several environments of one VecGame served over the reference's ZMQ +
protobuf protocol at the same time (one port pair each, actions gathered into batched steps);
every agent-side protocol client must see exactly what the CPU oracle computes."""
import threading

import numpy as np
import pytest

from helpers import WireClient
from test_wire import free_ports
from xroute_env_b200.instances import ispd18_geometry, make_batch

pytestmark = pytest.mark.gpu


def test_vecgame_served_over_the_wire():
    pytest.importorskip("zmq")
    from oracle.oracle import OracleEnv
    from xroute_env_b200 import VecGame
    from xroute_env_b200.wire import BatchDispatcher, SimulatorServer, VecGameBackend, request_to_data
    from xroute_env_b200 import build_3Dgrid
    g = ispd18_geometry(20, 18, 5)
    insts = make_batch(g, 3, 6, seed=77)
    vg = VecGame(g, insts, device=0)
    vg.reset()
    disp = BatchDispatcher(vg, max_wait_s=0.002)
    ports = free_ports(6)
    servers = [SimulatorServer(VecGameBackend(vg, e, disp), data_port=ports[2 * e], ctrl_port=ports[2 * e + 1],
                               dense=(e == 0)).start() for e in range(3)]
    errors = []

    def agent(e):
        try:
            cli = WireClient(ports[2 * e], ports[2 * e + 1])
            for episode in range(2):
                orc = OracleEnv(g, insts[e])
                msg = cli.reset()
                assert msg["metrics"] == [0, 0, 0] and [n + 1 for n in msg["nets"]] == insts[e].net_ids
                routed = set()
                for net in np.random.default_rng(e + 10 * episode).permutation(insts[e].net_ids):
                    msg = cli.step(int(net))
                    routed.add(int(net))
                    m = orc.step(int(net))
                    assert msg["metrics"] == [m["violation"], m["wirelength"], m["via"]], (e, net)
                    assert [n + 1 for n in msg["nets"]] == orc.remaining()
                    # the observation the reference would build from this message == the oracle's
                    obs = build_3Dgrid(request_to_data(msg), routed)[0]
                    assert np.array_equal(obs.numpy(), orc.obs()), (e, net)
                assert msg["is_done"]
            cli.close()
        except Exception as ex:      # surfaced in the main thread
            errors.append((e, repr(ex)))

    ts = [threading.Thread(target=agent, args=(e,)) for e in range(3)]
    for t in ts: t.start()
    for t in ts: t.join(timeout=120)
    for s in servers: s.close()
    disp.close()
    vg.close()
    assert not errors, errors
    assert sum(s.steps for s in servers) == 2 * sum(len(i.net_ids) for i in insts)
    assert disp.batches <= sum(s.steps for s in servers)


def test_serve_cli_plays_an_episode():
    """`python -m xroute_env_b200.serve` (what replaces launch_training.py + the simulator container): an agent-side
    protocol client connects to environment 1 of a 2-environment server and plays a whole episode."""
    pytest.importorskip("zmq")
    import os
    import subprocess
    import sys
    import time
    from oracle.oracle import OracleEnv
    from xroute_env_b200.instances import preset_geometry
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dp, cp = free_ports(2)
    dp, cp = dp - 1, cp - 1                                   # environment 1 listens on base + 1
    proc = subprocess.Popen([sys.executable, "-m", "xroute_env_b200.serve", "--preset", "T1-1x1", "--envs", "2", "--nets", "5",
                             "--seed", "3", "--data-port", str(dp), "--ctrl-port", str(cp), "--seconds", "60", "--sparse"],
                            cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        line = proc.stdout.readline()
        assert line.startswith("serving 2 environment(s) 25x26x9"), line
        geom = preset_geometry("T1-1x1")
        inst = make_batch(geom, 2, 5, 3)[1]
        orc = OracleEnv(geom, inst)
        cli = WireClient(dp + 1, cp + 1)
        msg = cli.reset()
        assert [n + 1 for n in msg["nets"]] == inst.net_ids
        for net in np.random.default_rng(0).permutation(inst.net_ids):
            msg = cli.step(int(net))
            m = orc.step(int(net))
            assert msg["metrics"] == [m["violation"], m["wirelength"], m["via"]]
        assert msg["is_done"]
        cli.close()
    finally:
        proc.terminate()
        try:
            proc.wait(timeout=10)
        except Exception:
            proc.kill()
