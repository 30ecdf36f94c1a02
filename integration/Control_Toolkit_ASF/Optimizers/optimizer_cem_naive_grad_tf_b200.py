"""Drop-in shim for the reference's glob-based optimizer discovery (others/globals_and_utils.py:103-133): copy into the
application's Control_Toolkit_ASF/Optimizers/ and select `optimizer: cem-naive-grad-tf-b200` in config_controllers.yml."""
from control_toolkit_b200.Optimizers.optimizer_cem_naive_grad_tf import optimizer_cem_naive_grad_tf


class optimizer_cem_naive_grad_tf_b200(optimizer_cem_naive_grad_tf):
    pass
