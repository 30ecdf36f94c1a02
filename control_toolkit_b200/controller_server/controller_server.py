"""controller_server -- the reference's ZeroMQ serving edge in front of the B200 backend
(reference controller_server/controller_server.py:55-86; its client is the reference's own Controllers/controller_remote.py:54-108,
which keeps working unchanged).

Wire protocol, as in the reference: a ROUTER socket receives ``[identity, payload]`` or ``[identity, b"", payload]`` frames; the
payload is UTF-8 JSON ``{"rid": int, "state": [...], "time": float | null, "updated_attributes": {...}}``; the reply to the same
identity is ``{"rid": rid, "Q": float | [...]}``.  A request that raises is logged and gets NO reply (the client's 50 ms timeout
and rid check handle it, controller_remote.py:80-101); frames with any other shape are skipped.

What differs: the controller behind the socket is ``control_toolkit_b200.Controllers.controller_mpc`` (one CUDA launch per tick),
chosen by command-line options instead of the reference's Tk dialog (controller_server/gui.py is out of scope); requests that
are already queued when a tick finishes are drained in arrival order without going back to ``poll`` (a tick is 30-60 us, the
JSON and socket work around it dominates, so the socket is never left idle between queued requests).

    python -m control_toolkit_b200.controller_server.controller_server --optimizer mppi --endpoint tcp://*:5555
"""
from __future__ import annotations

import argparse
import json
import sys
from typing import Optional

import numpy as np

ENDPOINT = "tcp://*:5555"  # the reference's hardcoded endpoint (:16)

# the attributes the reference server initialises (:19-26)
INITIAL_ENVIRONMENT_ATTRIBUTES = {"target_position": 0.0, "target_equilibrium": 0.0, "m_pole": 0.0, "L": 0.0, "Q_ccrc": 0.0,
                                  "Q_applied_-1": 0.0}


def _q_payload(Q):
    """reference :74-79: arrays as lists, scalars as floats, lists / tuples as they are."""
    if isinstance(Q, np.ndarray):
        return Q.tolist()
    return float(Q) if not isinstance(Q, (list, tuple)) else Q


def handle_request(ctrl, payload: bytes) -> bytes:
    """One request -> one reply payload (raises on a malformed request or a controller error, like the reference's try block)."""
    req = json.loads(payload.decode("utf-8"))
    rid = req["rid"]
    s = np.asarray(req["state"], dtype=np.float32)
    t = req.get("time")
    upd = req.get("updated_attributes", {}) or {}
    Q = ctrl.step(s, t, upd)
    return json.dumps({"rid": rid, "Q": _q_payload(Q)}).encode("utf-8")


def serve(ctrl, endpoint: str = ENDPOINT, max_requests: Optional[int] = None, poll_ms: int = 100, stop=None, context=None) -> int:
    """Serve ``ctrl.step`` on a ROUTER socket until ``max_requests`` frames were handled or ``stop()`` is true.
    Returns the number of replies sent."""
    import zmq

    ctx = context or zmq.Context.instance()
    sock = ctx.socket(zmq.ROUTER)
    sock.bind(endpoint)
    poller = zmq.Poller()
    poller.register(sock, zmq.POLLIN)
    handled = replied = 0
    try:
        while (max_requests is None or handled < max_requests) and not (stop is not None and stop()):
            if not poller.poll(poll_ms):
                continue
            while max_requests is None or handled < max_requests:  # drain what is queued, in arrival order
                try:
                    parts = sock.recv_multipart(flags=zmq.NOBLOCK)
                except zmq.Again:
                    break
                handled += 1
                if len(parts) == 2:
                    identity, payload = parts
                elif len(parts) == 3 and parts[1] == b"":
                    identity, _empty, payload = parts
                else:
                    continue  # unexpected framing: skipped, as in the reference (:62-69)
                try:
                    reply = handle_request(ctrl, payload)
                except Exception as e:  # noqa: BLE001  reference :84-86: log, send nothing back
                    print(f"[server] controller exception - no reply sent: {e}", file=sys.stderr)
                    continue
                sock.send_multipart([identity, reply])
                replied += 1
    finally:
        sock.close(linger=0)
    return replied


def build_controller(optimizer: str, predictor: str = "ODE", cost: str = "default", config_optimizers: Optional[dict] = None):
    """The controller the reference server builds at :38-49, on the B200 backend."""
    from ..Controllers.controller_mpc import controller_mpc

    ctrl = controller_mpc(
        environment_name="CartPole",
        control_limits=(np.array([-1.0], np.float32), np.array([1.0], np.float32)),
        initial_environment_attributes=dict(INITIAL_ENVIRONMENT_ATTRIBUTES),
        config_controller=dict(optimizer=optimizer, predictor_specification=predictor, cost_function_specification=cost,
                               controller_logging=False, calculate_optimal_trajectory=False) if config_optimizers is not None else None,
        config_optimizers=config_optimizers,
    )
    ctrl.configure(optimizer_name=optimizer, predictor_specification=predictor)
    return ctrl


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--optimizer", default=None, help="optimizer key of config_optimizers.yml (default: the one in config_controllers.yml)")
    ap.add_argument("--predictor", default="ODE")
    ap.add_argument("--endpoint", default=ENDPOINT)
    args = ap.parse_args(argv)
    ctrl = build_controller(args.optimizer, args.predictor)
    print(f"[server] controller: mpc   optimizer: {ctrl.optimizer.optimizer_name}   listening on {args.endpoint}")
    serve(ctrl, args.endpoint)


if __name__ == "__main__":
    main()
