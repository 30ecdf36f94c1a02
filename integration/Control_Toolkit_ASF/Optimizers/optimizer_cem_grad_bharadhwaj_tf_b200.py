"""Drop-in shim for the reference's glob-based optimizer discovery (others/globals_and_utils.py:103-133): copy into the
application's Control_Toolkit_ASF/Optimizers/ and select `optimizer: cem-grad-bharadhwaj-tf-b200` in config_controllers.yml."""
from control_toolkit_b200.Optimizers.optimizer_cem_grad_bharadhwaj_tf import optimizer_cem_grad_bharadhwaj_tf


class optimizer_cem_grad_bharadhwaj_tf_b200(optimizer_cem_grad_bharadhwaj_tf):
    pass
