"""Drop-in shim: copy (or symlink) into the application's ``Control_Toolkit_ASF/Optimizers/`` so the reference's own
``import_optimizer_by_name("mppi-b200")`` (others/globals_and_utils.py:103-133) finds the B200 backend.  The file name /
class name rule of the reference requires a unique ``optimizer_<name>.py`` holding ``class optimizer_<name>``."""
from control_toolkit_b200.Optimizers.optimizer_mppi import optimizer_mppi


class optimizer_mppi_b200(optimizer_mppi):
    pass
