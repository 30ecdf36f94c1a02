"""Drop-in shim for the reference's glob-based optimizer discovery (others/globals_and_utils.py:103-133): copy into the
application's Control_Toolkit_ASF/Optimizers/ and select `optimizer: gradient-tf-b200` in config_controllers.yml."""
from control_toolkit_b200.Optimizers.optimizer_gradient_tf import optimizer_gradient_tf


class optimizer_gradient_tf_b200(optimizer_gradient_tf):
    pass
