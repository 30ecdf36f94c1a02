"""The ZeroMQ serving edge (reference controller_server/controller_server.py:55-86) with the reference's client
protocol (Controllers/controller_remote.py:66-108): framing, rid echo, payload conversion, silent drop on errors.  CPU part: a
stand-in controller; GPU part: the real controller_mpc behind the socket against a local one."""
import json
import socket
import threading

import numpy as np
import pytest

zmq = pytest.importorskip("zmq")

from control_toolkit_b200.controller_server import controller_server as cs  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Echo:
    """step(s, time, updated_attributes) like controller_mpc.step; raises on a marked state."""
    def __init__(self):
        self.calls = []

    def step(self, s, time=None, updated_attributes={}):
        self.calls.append((np.asarray(s).copy(), time, dict(updated_attributes)))
        if s[0] < -100:
            raise RuntimeError("boom")
        if s[0] > 100:
            return np.array([s[1], s[2]], np.float32)  # vector-valued control
        return np.float32(s[1] * 2.0)


def _start(ctrl, n):
    port = _free_port()
    ctx = zmq.Context()
    th = threading.Thread(target=cs.serve, kwargs=dict(ctrl=ctrl, endpoint=f"tcp://127.0.0.1:{port}", max_requests=n, context=ctx), daemon=True)
    th.start()
    return ctx, th, port


def test_protocol_framing_rid_and_error_behaviour():
    ctrl = _Echo()
    ctx, th, port = _start(ctrl, 6)
    dealer = ctx.socket(zmq.DEALER)  # the reference client's socket type: frames arrive as [identity, payload]
    dealer.setsockopt(zmq.RCVTIMEO, 2000)
    dealer.connect(f"tcp://127.0.0.1:{port}")
    req = ctx.socket(zmq.DEALER)     # a client that sends the REQ-style empty delimiter itself: [identity, b"", payload]
    req.setsockopt(zmq.RCVTIMEO, 2000)
    req.connect(f"tcp://127.0.0.1:{port}")
    try:
        dealer.send_json({"rid": 7, "state": [0.0, 0.25, 0, 0, 0, 0], "time": 0.5, "updated_attributes": {"target_position": 0.1}})
        r = dealer.recv_json()
        assert r == {"rid": 7, "Q": 0.5}
        assert ctrl.calls[-1][1] == 0.5 and ctrl.calls[-1][2] == {"target_position": 0.1}
        dealer.send_json({"rid": 8, "state": [200.0, 1.0, 2.0, 0, 0, 0], "time": None})  # missing updated_attributes -> {}
        assert dealer.recv_json() == {"rid": 8, "Q": [1.0, 2.0]}
        req.send_multipart([b"", json.dumps({"rid": 1, "state": [0.0, -1.0, 0, 0, 0, 0], "time": None, "updated_attributes": {}}).encode()])
        assert req.recv_json() == {"rid": 1, "Q": -2.0}  # the reply is [identity, payload] without a delimiter, as the reference sends it (:82)
        dealer.send(b"this is not json")                                        # malformed: no reply
        dealer.send_json({"rid": 9, "state": [-200.0, 0, 0, 0, 0, 0]})           # controller raises: no reply
        dealer.send_json({"rid": 10, "state": [0.0, 3.0, 0, 0, 0, 0]})           # ... and the server keeps serving
        assert dealer.recv_json() == {"rid": 10, "Q": 6.0}
    finally:
        dealer.close(linger=0)
        req.close(linger=0)
        th.join(timeout=5)
        ctx.term()
    assert not th.is_alive()


def test_handle_request_payloads():
    ctrl = _Echo()
    out = json.loads(cs.handle_request(ctrl, json.dumps({"rid": 3, "state": [0, 1.5, 0, 0, 0, 0]}).encode()))
    assert out == {"rid": 3, "Q": 3.0}
    with pytest.raises(KeyError):
        cs.handle_request(ctrl, json.dumps({"state": [0, 1, 0, 0, 0, 0]}).encode())  # no rid -> dropped by serve()


@pytest.mark.gpu
def test_remote_controller_equals_local_controller():
    """The real backend behind the socket: a client stepping through the server gets the controls a local controller computes
    (same seed -> same in-kernel Philox noise), including a live target change sent as updated_attributes."""
    from oracle import spec
    cfg = {"mppi": dict(seed=42, mpc_horizon=50, mpc_timestep=0.02, num_rollouts=2000, cc_weight=1.0, R=1.0, LBD=100.0, NU=1000.0,
                        SQRTRHOINV=0.03, period_interpolation_inducing_points=10)}
    remote = cs.build_controller("mppi", config_optimizers=cfg)
    local = cs.build_controller("mppi", config_optimizers=cfg)
    states = spec.synthetic_states(5, seed=3)
    ctx, th, port = _start(remote, len(states))
    sock = ctx.socket(zmq.DEALER)
    sock.setsockopt(zmq.RCVTIMEO, 20000)
    sock.connect(f"tcp://127.0.0.1:{port}")
    try:
        for i, s in enumerate(states):
            upd = {"target_position": 0.05, "target_equilibrium": 1.0} if i >= 2 else {"target_equilibrium": 1.0}
            sock.send_json({"rid": i, "state": s.tolist(), "time": 0.02 * i, "updated_attributes": upd})
            r = sock.recv_json()
            u = local.step(s, 0.02 * i, upd)
            assert r["rid"] == i and np.float32(r["Q"]) == np.float32(u), (i, r, u)
    finally:
        sock.close(linger=0)
        th.join(timeout=5)
        ctx.term()


class _BatchOpt:
    """Stand-in for an MPPI optimizer built with num_clients = 3: u = 10 * slot + state[1]; records every round."""
    num_clients = 3

    def __init__(self):
        self.rounds, self.resets = [], []

    def reset_client(self, c):
        self.resets.append(c)

    def step_batch(self, states, active):
        self.rounds.append((np.array(states), np.array(active)))
        return np.array([10.0 * c + states[c][1] if active[c] else np.nan for c in range(self.num_clients)], np.float32)


class _BatchCtrl:
    def __init__(self):
        self.optimizer = _BatchOpt()
        self.updates = []

    def update_attributes(self, upd):
        self.updates.append(dict(upd))


def test_batched_server_rounds_slots_and_ordering():
    """serve_batched: one slot per ZeroMQ identity in order of first appearance, at most one request of a client per
    round (a client's second queued request waits for the next round), every reply carries the client's own rid and control, a
    client beyond the slot count gets no reply, malformed requests are dropped."""
    ctrl = _BatchCtrl()
    port = _free_port()
    ctx = zmq.Context()
    n_frames = 9
    th = threading.Thread(target=cs.serve_batched, kwargs=dict(ctrl=ctrl, endpoint=f"tcp://127.0.0.1:{port}", max_requests=n_frames, context=ctx),
                          daemon=True)
    th.start()
    socks = []
    for _ in range(4):
        s = ctx.socket(zmq.DEALER)
        s.setsockopt(zmq.RCVTIMEO, 3000)
        s.connect(f"tcp://127.0.0.1:{port}")
        socks.append(s)
    try:
        # client 0 first (slot 0), then 1, 2; their states carry a marker in state[1]
        for i in range(3):
            socks[i].send_json({"rid": 100 + i, "state": [0.0, 0.5 + i, 0, 0, 0, 0], "time": 0.0, "updated_attributes": {"target_position": 0.01 * i} if i == 1 else {}})
            r = socks[i].recv_json()
            assert r == {"rid": 100 + i, "Q": pytest.approx(10.0 * i + 0.5 + i)}
        assert ctrl.optimizer.resets == [0, 1, 2] and ctrl.updates == [{"target_position": 0.01}]
        # two queued requests of client 1 and one of client 2: answered in order, client 1's second one in a later round
        socks[1].send_json({"rid": 1, "state": [0.0, 1.0, 0, 0, 0, 0]})
        socks[1].send_json({"rid": 2, "state": [0.0, 2.0, 0, 0, 0, 0]})
        socks[2].send_json({"rid": 3, "state": [0.0, 3.0, 0, 0, 0, 0]})
        assert socks[1].recv_json() == {"rid": 1, "Q": pytest.approx(11.0)}
        assert socks[1].recv_json() == {"rid": 2, "Q": pytest.approx(12.0)}
        assert socks[2].recv_json() == {"rid": 3, "Q": pytest.approx(23.0)}
        socks[3].send_json({"rid": 9, "state": [0.0, 9.0, 0, 0, 0, 0]})  # a fourth identity: no slot, no reply
        socks[0].send(b"not json")                                        # malformed: dropped
        socks[0].send_json({"rid": 4, "state": [0.0, 4.0, 0, 0, 0, 0]})
        assert socks[0].recv_json() == {"rid": 4, "Q": pytest.approx(4.0)}
        with pytest.raises(zmq.Again):
            socks[3].setsockopt(zmq.RCVTIMEO, 200)
            socks[3].recv_json()
    finally:
        for s in socks:
            s.close(linger=0)
        th.join(timeout=5)
        ctx.term()
    assert not th.is_alive()
    for states, active in ctrl.optimizer.rounds:  # a round never holds two requests of one client
        assert active.sum() >= 1


@pytest.mark.gpu
def test_step_batch_equals_independent_controllers():
    """B clients' ticks in ONE launch (ctk_step_batch, grid.y = client slot) are bit-identical to B controllers of their
    own (same configuration and seed -> same Philox streams per tick count), also when clients skip rounds and when a slot is handed
    to a new client."""
    from oracle import spec
    B, T = 4, 6
    base = dict(seed=42, mpc_horizon=50, mpc_timestep=0.02, num_rollouts=2000, cc_weight=1.0, R=1.0, LBD=100.0, NU=1000.0,
                SQRTRHOINV=0.03, period_interpolation_inducing_points=10)
    batched = cs.build_controller("mppi", config_optimizers={"mppi": dict(base, num_clients=B)})
    locals_ = [cs.build_controller("mppi", config_optimizers={"mppi": dict(base)}) for _ in range(B)]
    states = spec.synthetic_states(B * T, seed=5).reshape(T, B, 6)
    rng = np.random.default_rng(0)
    n0 = batched.optimizer.gpu_launches
    rounds = 0
    for t in range(T):
        active = rng.random(B) < 0.7 if t not in (0, T - 1) else np.ones(B, bool)
        if not active.any():
            continue
        if t == 3:  # slot 2 is taken over by a new client
            batched.optimizer.reset_client(2)
            locals_[2] = cs.build_controller("mppi", config_optimizers={"mppi": dict(base)})
        u = batched.optimizer.step_batch(states[t], active)
        rounds += 1
        for c in range(B):
            if active[c]:
                assert np.float32(u[c]) == np.float32(locals_[c].step(states[t, c], 0.02 * t)), (t, c)
            else:
                assert np.isnan(u[c])
    assert batched.optimizer.gpu_launches - n0 == rounds  # ONE launch per round, whatever the number of active clients
    assert "mppi_ode_batch_kernel" in batched.optimizer.last_kernel
    from control_toolkit_b200 import _lib as L
    u_nom = batched.optimizer._get_state(L.STATE_U_NOM, (B, 50))
    for c in range(B):
        np.testing.assert_array_equal(u_nom[c], locals_[c].optimizer.u_nom.ravel())


@pytest.mark.gpu
def test_batched_server_equals_local_controllers():
    """Three remote clients behind serve_batched get the controls three local controllers compute."""
    from oracle import spec
    B, T = 3, 4
    base = dict(seed=42, mpc_horizon=50, mpc_timestep=0.02, num_rollouts=2000, cc_weight=1.0, R=1.0, LBD=100.0, NU=1000.0,
                SQRTRHOINV=0.03, period_interpolation_inducing_points=10)
    remote = cs.build_controller("mppi", config_optimizers={"mppi": dict(base, num_clients=B)})
    locals_ = [cs.build_controller("mppi", config_optimizers={"mppi": dict(base)}) for _ in range(B)]
    states = spec.synthetic_states(B * T, seed=6).reshape(T, B, 6)
    port = _free_port()
    ctx = zmq.Context()
    th = threading.Thread(target=cs.serve_batched, kwargs=dict(ctrl=remote, endpoint=f"tcp://127.0.0.1:{port}", max_requests=B * T, context=ctx,
                                                               gather_ms=20.0), daemon=True)
    th.start()
    socks = []
    for _ in range(B):
        s = ctx.socket(zmq.DEALER)
        s.setsockopt(zmq.RCVTIMEO, 20000)
        s.connect(f"tcp://127.0.0.1:{port}")
        socks.append(s)
    try:
        socks[0].send_json({"rid": -1 + 1000, "state": states[0, 0].tolist(), "time": 0.0, "updated_attributes": {"target_equilibrium": 1.0}})
        r0 = socks[0].recv_json()  # slot order = order of first appearance: client 0 first
        for c in range(1, B):
            socks[c].send_json({"rid": 1000 + c, "state": states[0, c].tolist(), "time": 0.0})
        rs = [r0] + [socks[c].recv_json() for c in range(1, B)]
        for c in range(B):
            locals_[c].update_attributes({"target_equilibrium": 1.0})
            assert np.float32(rs[c]["Q"]) == np.float32(locals_[c].step(states[0, c], 0.0)), c
        for t in range(1, T):
            for c in range(B):
                socks[c].send_json({"rid": 10 * t + c, "state": states[t, c].tolist(), "time": 0.02 * t})
            for c in range(B):
                r = socks[c].recv_json()
                assert r["rid"] == 10 * t + c and np.float32(r["Q"]) == np.float32(locals_[c].step(states[t, c], 0.02 * t)), (t, c)
    finally:
        for s in socks:
            s.close(linger=0)
        th.join(timeout=5)
        ctx.term()
