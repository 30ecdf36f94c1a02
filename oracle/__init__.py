"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (torch-CPU fp32 / numpy) of the batched MPC rollout hot path of
SensorsINI/Control_Toolkit (MPPI / CEM / RPGD + the predictor and cost they drive).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import anything from here, and only as the *checker* or the
*reported CPU baseline*.  Nothing under ``control_toolkit_b200/`` imports this package.

Pinning status
--------------
* ``oracle.mppi`` / ``oracle.rpgd``  : pinned against the reference's UNMODIFIED
  ``Optimizers/optimizer_mppi.py`` / ``optimizer_rpgd.py`` executed in this container through
  the shims in ``oracle/refharness`` (fixtures in ``tests/golden``, generator
  ``oracle/gen_golden.py``).
* ``oracle.cem``                   : pinned against the reference's UNMODIFIED
  ``Optimizers/optimizer_cem_tf.py`` executed through a TensorFlow-*API* shim over torch-CPU
  (TensorFlow itself is not installable here), i.e. control flow and op sequence are the
  reference file's; TF's own numerics are restated from its documented semantics.
* ``oracle.spec`` (CartPole ODE, cost functions, MLP predictor) : **parity unpinned** --
  the reference does not contain this arithmetic (it lives in the un-vendored, un-pinned
  SI_Toolkit / CartPoleSimulation repos).  It is this build's written spec.
"""
