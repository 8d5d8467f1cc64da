"""Wire-level drop-in: the *simulator side* of the reference's
ZMQ + protobuf protocol, served by the GPU environment.

The reference agent never links the router: ``Game`` (``/root/reference/baseline/baseline_utils.py:383-481``)
binds a REP socket on ``tcp://*:5556``, asks a launcher on ``tcp://127.0.0.1:6667`` to (re)start
the simulator with ``b'initial'`` (``:455-459``; launcher loop ``examples/launch_training.py:89-102``),
and then answers every ``Message{Request}`` the simulator sends with a ``Message{Response{net_index}}``
(``:405-419``; schema ``baseline/openroad_api/proto/net_ordering.proto:1-56``).  ``SimulatorServer``
plays launcher + simulator, so the UNMODIFIED reference ``Game``, ``train_PPO.py``/``train_DQN.py`` and
the evaluation servers run against this library without a line changed.

* ``encode_request`` / ``encode_response`` / ``decode_message``: a hand-written proto3 codec for exactly
  this schema (numpy-vectorised node stream; no protoc, no generated module).
* ``SimulatorServer``: control socket (REP) + data socket (REQ) for one environment.
* ``VecGameBackend``: one environment of a ``VecGame`` batch as the simulator state; ``BatchDispatcher``
  gathers the actions of many concurrently served environments into single batched ``xr_step`` calls.
"""
from __future__ import annotations

import threading
import time

import numpy as np

BLOCKAGE, NORMAL, ACCESS = 0, 1, 2          # net_ordering.proto:5-9


# ------------------------------------------------------------------ proto3 codec
def _varint_bytes(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _zz(v: int) -> int:
    return (v << 1) ^ (v >> 63)


def _read_varint(b, i):
    r = s = 0
    while True:
        c = b[i]; i += 1
        r |= (c & 0x7F) << s; s += 7
        if not c & 0x80:
            return r, i


def _unzz(v: int) -> int:
    return (v >> 1) ^ -(v & 1)


def _fields(b):
    i, n = 0, len(b)
    while i < n:
        key, i = _read_varint(b, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _read_varint(b, i)
        elif wt == 2:
            ln, i = _read_varint(b, i); v = b[i:i + ln]; i += ln
        elif wt == 1:
            v = b[i:i + 8]; i += 8
        elif wt == 5:
            v = b[i:i + 4]; i += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fn, wt, v


def _encode_nodes(cols) -> bytes:
    """Vectorised ``repeated Node nodes = 4``: ``cols`` = 10 int64 arrays in field order
    (maze xyz, point xyz, type, is_used, net, pin); sint32 fields zigzag-encoded, proto3
    default (zero) values omitted, every node framed as tag 0x22 + length + body."""
    n = len(cols[0])
    if n == 0:
        return b""
    sint = (True, True, True, True, True, True, False, False, True, True)
    vals, lens = [], []
    for c, s in zip(cols, sint):
        v = np.asarray(c, np.int64)
        u = ((v << 1) ^ (v >> 63)).astype(np.uint64) if s else v.astype(np.uint64)
        ln = np.ones(n, np.int64)
        for k in range(1, 5):
            ln += (u >= (1 << (7 * k))).astype(np.int64)
        ln[u == 0] = 0                                       # omitted
        vals.append(u); lens.append(ln)
    size = np.zeros(n, np.int64)
    for ln in lens:
        size += ln + (ln > 0)
    assert int(size.max()) < 128
    start = np.concatenate([[0], np.cumsum(size + 2)[:-1]])
    out = np.zeros(int((size + 2).sum()), np.uint8)
    out[start] = 0x22
    out[start + 1] = size.astype(np.uint8)
    pos = start + 2
    for f, (u, ln) in enumerate(zip(vals, lens), 1):
        m = ln > 0
        out[pos[m]] = (f << 3)
        for k in range(5):
            mk = ln > k
            if not mk.any():
                break
            byte = ((u[mk] >> np.uint64(7 * k)) & np.uint64(0x7F)).astype(np.uint8)
            byte |= ((ln[mk] > k + 1).astype(np.uint8) << 7)
            out[pos[mk] + 1 + k] = byte
        pos = pos + ln + m
    return out.tobytes()


def encode_request(dims, nodes, metrics, nets, is_done=False) -> bytes:
    """``Message{request}``.  ``nodes``: dict of equally long arrays ``maze`` [n,3], ``point`` [n,3],
    ``type``, ``is_used``, ``net``, ``pin`` (0-based ids, -1 = none: net_ordering.proto:24-26);
    ``metrics``: cumulative (violation, wire_length, via); ``nets``: 0-based ids still to route."""
    body = bytearray()
    for f, v in zip((1, 2, 3), dims):
        if v:
            body += bytes([f << 3]) + _varint_bytes(int(v))
    maze, point = np.asarray(nodes["maze"], np.int64).reshape(-1, 3), np.asarray(nodes["point"], np.int64).reshape(-1, 3)
    body += _encode_nodes([maze[:, 0], maze[:, 1], maze[:, 2], point[:, 0], point[:, 1], point[:, 2],
                           nodes["type"], np.asarray(nodes["is_used"]).astype(np.int64), nodes["net"], nodes["pin"]])
    for f, v in zip((5, 6, 7), metrics):
        if v:
            body += bytes([f << 3]) + _varint_bytes(int(v))
    if is_done:
        body += bytes([8 << 3, 1])
    if len(nets):
        packed = b"".join(_varint_bytes(int(v)) for v in nets)
        body += bytes([(9 << 3) | 2]) + _varint_bytes(len(packed)) + packed
    return bytes([(1 << 3) | 2]) + _varint_bytes(len(body)) + bytes(body)


def encode_response(net_index: int) -> bytes:
    body = bytes([1 << 3]) + _varint_bytes(_zz(int(net_index))) if net_index else b""
    return bytes([(2 << 3) | 2]) + _varint_bytes(len(body)) + body


def decode_message(raw: bytes):
    """-> ``("request", dict)`` or ``("response", net_index)``; the request dict mirrors
    ``encode_request``'s arguments (node columns as int64 arrays)."""
    for fn, wt, v in _fields(raw):
        if fn == 2 and wt == 2:
            idx = 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    idx = _unzz(v2)
            return "response", idx
        if fn == 1 and wt == 2:
            req = {"dims": [0, 0, 0], "metrics": [0, 0, 0], "is_done": False, "nets": []}
            rows = []
            for f2, wt2, v2 in _fields(v):
                if f2 in (1, 2, 3):
                    req["dims"][f2 - 1] = v2
                elif f2 == 4:
                    r = [0] * 10
                    for f3, _, v3 in _fields(v2):
                        if 1 <= f3 <= 10:
                            r[f3 - 1] = v3 if f3 in (7, 8) else _unzz(v3)
                    rows.append(r)
                elif f2 in (5, 6, 7):
                    req["metrics"][f2 - 5] = v2
                elif f2 == 8:
                    req["is_done"] = bool(v2)
                elif f2 == 9:
                    if wt2 == 2:
                        i = 0
                        while i < len(v2):
                            x, i = _read_varint(v2, i)
                            req["nets"].append(x)
                    else:
                        req["nets"].append(v2)
            a = np.asarray(rows, np.int64).reshape(-1, 10)
            req["nodes"] = {"maze": a[:, 0:3], "point": a[:, 3:6], "type": a[:, 6], "is_used": a[:, 7],
                            "net": a[:, 8], "pin": a[:, 9]}
            return "request", req
    raise ValueError("empty Message")


def request_to_data(req) -> list:
    """The ``data`` list ``handle_messange`` builds from a Request (``baseline_utils.py:16-40``):
    ``[[X,Y,Z], nodes, [vio, wl, via], nets+1]``, node ``[[mx,my,mz],[px,py,pz],[used, Net, Pin]]``
    with ``Net`` = net+1 for ACCESS, -1 for BLOCKAGE, 0 otherwise and ``Pin`` = pin+1 for ACCESS else -1."""
    nd = req["nodes"]
    acc = nd["type"] == ACCESS
    net = np.where(acc, nd["net"] + 1, np.where(nd["type"] == BLOCKAGE, -1, 0))
    pin = np.where(acc, nd["pin"] + 1, -1)
    nodes = [[[int(m[0]), int(m[1]), int(m[2])], [int(p[0]), int(p[1]), int(p[2])], [int(u), int(n), int(q)]]
             for m, p, u, n, q in zip(nd["maze"], nd["point"], nd["is_used"], net, pin)]
    return [[int(v) for v in req["dims"]], nodes, [int(v) for v in req["metrics"]], [int(v) + 1 for v in req["nets"]]]


# ------------------------------------------------------------------ node stream of a region state
def region_nodes(geom, inst, usage, dense: bool = True) -> dict:
    """Node columns of a region in occupancy state ``usage`` (uint8 [Z,Y,X]): every cell (``dense``)
    or only blockages, access points and used cells.  Raster order x, then y, then z fastest --
    the order is irrelevant to the reference's decoder."""
    X, Y, Z = geom.X, geom.Y, geom.Z
    typ = np.full((Z, Y, X), NORMAL, np.int64)
    net = np.full((Z, Y, X), -1, np.int64)
    pin = np.full((Z, Y, X), -1, np.int64)
    if len(inst.block_xyz):
        b = inst.block_xyz
        typ[b[:, 2], b[:, 1], b[:, 0]] = BLOCKAGE
    if len(inst.ap_xyz):
        a = inst.ap_xyz
        typ[a[:, 2], a[:, 1], a[:, 0]] = ACCESS
        net[a[:, 2], a[:, 1], a[:, 0]] = inst.ap_net - 1
        pin[a[:, 2], a[:, 1], a[:, 0]] = inst.ap_pin - 1
    used = (np.asarray(usage).reshape(Z, Y, X) > 0) | (typ == BLOCKAGE)
    zz, yy, xx = np.meshgrid(np.arange(Z), np.arange(Y), np.arange(X), indexing="ij")
    keep = np.ones((Z, Y, X), bool) if dense else ((typ != NORMAL) | used)
    order = np.lexsort((zz[keep], yy[keep], xx[keep]))
    sel = lambda a: a[keep][order]
    mx, my, mz = sel(xx), sel(yy), sel(zz)
    return {"maze": np.stack([mx, my, mz], 1),
            "point": np.stack([geom.x_coords.astype(np.int64)[mx], geom.y_coords.astype(np.int64)[my], mz], 1),
            "type": sel(typ), "is_used": sel(used).astype(np.int64), "net": sel(net), "pin": sel(pin)}


# ------------------------------------------------------------------ backends
class BatchDispatcher:
    """Micro-batching front end of a ``VecGame``: environments served on separate sockets submit
    their action and block; a dispatcher thread steps all pending actions in ONE batched
    ``xr_step`` (action 0 = idle for the others) -- the many-simulator-processes layout of the
    reference (one port per worker, ``baseline/A3C/discrete_A3C.py:246-247``) on one GPU handle."""

    def __init__(self, vg, max_wait_s: float = 0.0005):
        self.vg = vg
        self.max_wait_s = max_wait_s
        self.lock = threading.Lock()                 # the C handle is not thread-safe
        self.cv = threading.Condition()
        self.pending = {}                            # env -> action
        self.done = {}                               # env -> (cum metrics)
        self.batches = 0
        self.stop = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def submit(self, env: int, action: int):
        with self.cv:
            self.pending[env] = int(action)
            self.cv.notify_all()
            while env not in self.done and not self.stop:
                self.cv.wait(0.05)
            res = self.done.pop(env, None)
        if isinstance(res, Exception):               # the batched step failed: every caller of that batch sees why
            raise res
        return res

    def _run(self):
        while not self.stop:
            with self.cv:
                while not self.pending and not self.stop:
                    self.cv.wait(0.05)
                if self.stop:
                    return
            time.sleep(self.max_wait_s)              # let concurrent agents join the batch
            with self.cv:
                batch, self.pending = self.pending, {}
            acts = np.zeros(self.vg.n_envs, np.int32)
            for e, a in batch.items():
                acts[e] = a
            try:
                with self.lock:
                    self.vg.step(acts)
                    _, _, cum = self.vg.results_host()
                res = {e: [int(v) for v in cum[e][:3]] for e in batch}
            except Exception as exc:                 # illegal action, capacity error ...: hand it to every waiting caller
                res = {e: exc for e in batch}
            self.batches += 1
            with self.cv:
                self.done.update(res)
                self.cv.notify_all()

    def close(self):
        self.stop = True
        with self.cv:
            self.cv.notify_all()
        self.thread.join(timeout=2)


class VecGameBackend:
    """Environment ``env`` of a ``VecGame`` as simulator state (reset / step / snapshot)."""

    def __init__(self, vg, env: int = 0, dispatcher: BatchDispatcher | None = None):
        self.vg, self.env, self.dispatcher = vg, env, dispatcher
        self.geom, self.inst = vg.geom, vg.insts[env]
        self.cum = [0, 0, 0]
        self._lock = dispatcher.lock if dispatcher else threading.Lock()   # the C handle is not thread-safe

    def _locked(self):
        return self._lock

    def reset(self):
        with self._locked():
            self.vg.reset([self.env])
        self.cum = [0, 0, 0]

    def remaining(self):
        with self._locked():
            return sorted(self.vg.legal_set(self.env))

    def step(self, net_id: int):
        if self.dispatcher:
            self.cum = self.dispatcher.submit(self.env, net_id)
            return
        acts = np.zeros(self.vg.n_envs, np.int32)
        acts[self.env] = net_id
        with self._locked():
            self.vg.step(acts)
            _, _, cum = self.vg.results_host()
            self.cum = [int(v) for v in cum[self.env][:3]]

    def usage(self):
        with self._locked():
            return self.vg.state(self.env)[0]


# ------------------------------------------------------------------ server
class SimulatorServer:
    """Launcher + simulator side of the reference protocol for one environment.

    ``backend``: object with ``geom``, ``inst``, ``cum`` (cumulative violation/wirelength/via),
    ``reset()``, ``step(net_id 1-based)``, ``remaining() -> [net ids]`` and ``usage() -> uint8 [Z,Y,X]``.
    ``start()`` binds the control REP socket on ``ctrl_port`` (every ``b'initial'`` starts a new
    episode, aborting one in flight) and connects a REQ socket to the agent's ``data_port`` per
    episode: send the state, receive the chosen net (``net_index`` 0-based, -1 = stop,
    ``net_ordering.proto:47-49``), route it, repeat; the final state carries ``is_done`` and the agent
    acknowledges with ``b'\\0'`` (``baseline_utils.py:41-42``).
    """

    def __init__(self, backend, data_port: int = 5556, ctrl_port: int = 6667, host: str = "127.0.0.1",
                 dense: bool = True, evaluation: bool = False, max_episodes: int = 0):
        import zmq
        self.zmq = zmq
        self.backend, self.data_port, self.ctrl_port, self.host, self.dense = backend, data_port, ctrl_port, host, dense
        # evaluation mode: the protocol of the reference's inference servers (baseline/PPO/test_PPO.py:50-86,
        # baseline/DQN/test_DQN.py): a bare REP loop on the agent side, no control socket and no is_done message --
        # the simulator walks through its regions by itself (examples/launch_evaluation.py)
        self.evaluation, self.max_episodes = evaluation, max_episodes
        self.ctx = zmq.Context()
        self.episodes = 0
        self.steps = 0
        self.log = []                      # (episode, net_index received)
        self._stop = threading.Event()
        self._gen = 0
        self._threads = []

    def start(self):
        if self.evaluation:
            self.ctrl = None
            t = threading.Thread(target=self._evaluation_loop, daemon=True)
            t.start()
            self._threads.append(t)
            return self
        self.ctrl = self.ctx.socket(self.zmq.REP)
        self.ctrl.bind(f"tcp://{self.host}:{self.ctrl_port}")
        t = threading.Thread(target=self._ctrl_loop, daemon=True)
        t.start()
        self._threads.append(t)
        return self

    def _ctrl_loop(self):
        poller = self.zmq.Poller()
        poller.register(self.ctrl, self.zmq.POLLIN)
        while not self._stop.is_set():
            if not poller.poll(50):
                continue
            self.ctrl.recv()
            self.ctrl.send(b"\0")
            self._gen += 1                 # aborts an episode in flight (its generation is stale)
            t = threading.Thread(target=self._episode, args=(self._gen,), daemon=True)
            t.start()
            self._threads.append(t)

    def _evaluation_loop(self):
        zmq = self.zmq
        sock = self.ctx.socket(zmq.REQ)
        sock.setsockopt(zmq.LINGER, 0)
        sock.connect(f"tcp://{self.host}:{self.data_port}")
        poller = zmq.Poller()
        poller.register(sock, zmq.POLLIN)
        try:
            while not self._stop.is_set() and (self.max_episodes <= 0 or self.episodes < self.max_episodes):
                self.backend.reset()
                self.episodes += 1
                while not self._stop.is_set() and self.backend.remaining():
                    sock.send(self._snapshot(is_done=False))
                    while not poller.poll(50):
                        if self._stop.is_set():
                            return
                    kind, idx = decode_message(sock.recv())
                    self.log.append((self.episodes, idx))
                    if idx < 0:
                        break
                    self.backend.step(idx + 1)
                    self.steps += 1
        finally:
            sock.close(0)

    def _snapshot(self, is_done: bool) -> bytes:
        b = self.backend
        g = b.geom
        nets = [] if is_done else [n - 1 for n in b.remaining()]
        return encode_request((g.X, g.Y, g.Z), region_nodes(g, b.inst, b.usage(), self.dense), b.cum, nets, is_done)

    def _episode(self, gen: int):
        zmq = self.zmq
        sock = self.ctx.socket(zmq.REQ)
        sock.setsockopt(zmq.LINGER, 0)
        sock.connect(f"tcp://{self.host}:{self.data_port}")
        poller = zmq.Poller()
        poller.register(sock, zmq.POLLIN)
        try:
            self.backend.reset()
            self.episodes += 1
            while not self._stop.is_set() and gen == self._gen:
                remaining = self.backend.remaining()
                sock.send(self._snapshot(is_done=not remaining))
                while not poller.poll(50):
                    if self._stop.is_set() or gen != self._gen:
                        return
                raw = sock.recv()
                if raw == b"\0" or not remaining:
                    return                                   # the agent acknowledged is_done
                kind, idx = decode_message(raw)
                if kind != "response":
                    raise ValueError("expected a Response")
                self.log.append((self.episodes, idx))
                if idx < 0:
                    return                                   # -1: the agent stops the episode
                self.backend.step(idx + 1)
                self.steps += 1
        finally:
            sock.close(0)

    def close(self):
        self._stop.set()
        for t in self._threads:
            t.join(timeout=2)
        try:
            if self.ctrl is not None:
                self.ctrl.close(0)
        except Exception:
            pass
        self.ctx.term()
