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
