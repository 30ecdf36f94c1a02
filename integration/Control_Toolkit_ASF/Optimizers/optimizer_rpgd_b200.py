"""Drop-in shim: copy (or symlink) into the application's ``Control_Toolkit_ASF/Optimizers/`` so the reference's own
``import_optimizer_by_name("rpgd-b200")`` (others/globals_and_utils.py:103-133) finds the B200 backend.  The file name /
class name rule of the reference requires a unique ``optimizer_<name>.py`` holding ``class optimizer_<name>``."""
from control_toolkit_b200.Optimizers.optimizer_rpgd import optimizer_rpgd


class optimizer_rpgd_b200(optimizer_rpgd):
    pass
