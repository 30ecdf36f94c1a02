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
JSON and socket work around it dominates, so the socket is never left idle between queued requests); and ``serve_batched`` puts
several clients behind one GPU, one kernel launch per round of requests.

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


def serve_batched(ctrl, endpoint: str = ENDPOINT, max_requests: Optional[int] = None, poll_ms: int = 100, stop=None, context=None,
                  gather_ms: float = 0.0) -> int:
    """The same protocol with SEVERAL clients behind one GPU: ``ctrl`` is a controller_mpc whose MPPI optimizer was built with
    ``num_clients`` = B > 1.  Every ZeroMQ identity gets a client slot of its own (warm-start sequence, previous input, noise
    stream) the first time it is seen; the requests queued on the socket are drained, at most one per client is taken into a round,
    and the round's ticks run as ONE kernel launch (``optimizer.step_batch``: grid.y = client slot) -- the reference serves one
    ``ctrl.step`` per request (controller_server/controller_server.py:55-86).  A second request of a client that arrives within
    the same round waits for the next one, so every client sees its requests answered in order.  ``updated_attributes`` are applied
    to the shared controller (the clients of one server share the environment configuration).  ``gather_ms`` > 0 waits that long for
    further clients' requests before launching a round that does not yet hold one request of every known client.  More identities
    than slots: the extra clients get no reply (the reference client's 50 ms timeout handles it)."""
    import time as _time
    import zmq

    opt = ctrl.optimizer
    B = int(getattr(opt, "num_clients", 1))
    if B < 2 or not hasattr(opt, "step_batch"):
        raise ValueError("serve_batched needs an MPPI optimizer built with num_clients > 1")
    ctx = context or zmq.Context.instance()
    sock = ctx.socket(zmq.ROUTER)
    sock.bind(endpoint)
    poller = zmq.Poller()
    poller.register(sock, zmq.POLLIN)
    slots: dict[bytes, int] = {}
    backlog: list = []  # (identity, request) in arrival order
    handled = replied = 0

    def drain():
        nonlocal handled
        while max_requests is None or handled < max_requests:
            try:
                parts = sock.recv_multipart(flags=zmq.NOBLOCK)
            except zmq.Again:
                return
            handled += 1
            if len(parts) == 2:
                identity, payload = parts
            elif len(parts) == 3 and parts[1] == b"":
                identity, _empty, payload = parts
            else:
                continue
            try:
                req = json.loads(payload.decode("utf-8"))
                req["rid"], np.asarray(req["state"], dtype=np.float32)
            except Exception as e:  # noqa: BLE001
                print(f"[server] malformed request - no reply sent: {e}", file=sys.stderr)
                continue
            if identity not in slots:
                if len(slots) >= B:
                    print(f"[server] no free client slot for {identity!r} - no reply sent", file=sys.stderr)
                    continue
                slots[identity] = len(slots)
                opt.reset_client(slots[identity])
            backlog.append((identity, req))

    try:
        while (max_requests is None or handled < max_requests or backlog) and not (stop is not None and stop()):
            if not backlog and not poller.poll(poll_ms):
                continue
            drain()
            if gather_ms > 0 and len({i for i, _ in backlog}) < len(slots):
                t_end = _time.perf_counter() + gather_ms * 1e-3
                while _time.perf_counter() < t_end and len({i for i, _ in backlog}) < len(slots):
                    if poller.poll(max(int((t_end - _time.perf_counter()) * 1e3), 0)):
                        drain()
            # one round: the first queued request of every client
            round_reqs, rest, seen = [], [], set()
            for identity, req in backlog:
                if identity in seen:
                    rest.append((identity, req))
                else:
                    seen.add(identity)
                    round_reqs.append((identity, req))
            backlog = rest
            if not round_reqs:
                continue
            states = np.zeros((B, 6), np.float32)
            active = np.zeros(B, bool)
            for identity, req in round_reqs:
                c = slots[identity]
                states[c] = np.asarray(req["state"], dtype=np.float32)
                active[c] = True
                upd = req.get("updated_attributes", {}) or {}
                if upd:
                    ctrl.update_attributes(upd)
            try:
                u = opt.step_batch(states, active)
            except Exception as e:  # noqa: BLE001  reference :84-86: log, send nothing back
                print(f"[server] controller exception - no reply sent: {e}", file=sys.stderr)
                continue
            for identity, req in round_reqs:
                sock.send_multipart([identity, json.dumps({"rid": req["rid"], "Q": float(u[slots[identity]])}).encode("utf-8")])
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
    ap.add_argument("--clients", type=int, default=1, help="> 1: client slots of a batched MPPI server (one launch per round of requests)")
    args = ap.parse_args(argv)
    if args.clients > 1:
        from ..Controllers.controller_mpc import _load_yaml
        cfgs = _load_yaml("config_optimizers.yml")
        name = args.optimizer or "mppi"
        cfgs[name] = dict(cfgs[name], num_clients=args.clients)
        ctrl = build_controller(name, args.predictor, config_optimizers=cfgs)
        print(f"[server] controller: mpc   optimizer: {name}   {args.clients} client slots   listening on {args.endpoint}")
        serve_batched(ctrl, args.endpoint)
        return
    ctrl = build_controller(args.optimizer, args.predictor)
    print(f"[server] controller: mpc   optimizer: {ctrl.optimizer.optimizer_name}   listening on {args.endpoint}")
    serve(ctrl, args.endpoint)


if __name__ == "__main__":
    main()
