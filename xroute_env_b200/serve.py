"""Serve GPU environments to UNMODIFIED reference agents over their own protocol.

    python -m xroute_env_b200.serve --regions tests/golden/ispd18_test1_regions.npz --region t1_7x7_y79800
    python -m xroute_env_b200.serve --preset T1-1x1 --envs 8 --nets 16

Replaces ``python examples/launch_training.py`` + the OpenROAD container of the reference
(``/root/reference/examples/launch_training.py:89-102``, ``simulator/start_container``): environment e
listens for ``b'initial'`` on ``--ctrl-port + e`` (reference default 6667) and talks to the agent's REP
socket on ``--data-port + e`` (default 5556), so ``python train_PPO.py cpu`` / ``train_DQN.py`` /
``test_PPO.py`` of the reference connect as they would to the simulator.  All environments live in one
``VecGame``; concurrent agents' actions are stepped as one batch.
"""
from __future__ import annotations

import argparse
import time


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--preset", default="T1-1x1", help="synthetic grid preset (instances.PRESETS)")
    ap.add_argument("--regions", help="npz of extracted regions (xroute_env_b200.ispd.save_regions)")
    ap.add_argument("--region", help="region name inside --regions (default: the first)")
    ap.add_argument("--envs", type=int, default=1)
    ap.add_argument("--nets", type=int, default=16, help="nets per synthetic region")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--data-port", type=int, default=5556)
    ap.add_argument("--ctrl-port", type=int, default=6667)
    ap.add_argument("--sparse", action="store_true", help="send only blockage / access / used nodes")
    ap.add_argument("--evaluation", action="store_true",
                    help="drive an inference server (test_PPO.py / test_DQN.py): no control socket, no is_done message")
    ap.add_argument("--seconds", type=float, default=0.0, help="exit after this long (0 = serve until interrupted)")
    args = ap.parse_args(argv)

    from . import VecGame, make_batch, preset_geometry
    from .wire import BatchDispatcher, SimulatorServer, VecGameBackend
    if args.regions:
        from .ispd import load_regions
        regions = load_regions(args.regions)
        name = args.region or sorted(regions)[0]
        geom, inst = regions[name]
        insts = [inst] * args.envs
    else:
        geom = preset_geometry(args.preset)
        insts = make_batch(geom, args.envs, args.nets, args.seed)
    vg = VecGame(geom, insts, device=args.device)
    vg.reset()
    disp = BatchDispatcher(vg) if args.envs > 1 else None
    servers = [SimulatorServer(VecGameBackend(vg, e, disp), data_port=args.data_port + e, ctrl_port=args.ctrl_port + e,
                               dense=not args.sparse, evaluation=args.evaluation).start() for e in range(args.envs)]
    print(f"serving {args.envs} environment(s) {geom.X}x{geom.Y}x{geom.Z}: data ports {args.data_port}.., "
          f"control ports {args.ctrl_port}..", flush=True)
    t0 = time.time()
    try:
        while args.seconds <= 0 or time.time() - t0 < args.seconds:
            time.sleep(0.2)
    except KeyboardInterrupt:
        pass
    for s in servers:
        s.close()
    if disp:
        disp.close()
    print(f"episodes {sum(s.episodes for s in servers)}, steps {sum(s.steps for s in servers)}")
    vg.close()


if __name__ == "__main__":
    main()
