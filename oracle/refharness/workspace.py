"""TEST-ONLY: build the working directory the reference needs in order to be imported unmodified
(SURVEY.md section 8c): ``Control_Toolkit -> /root/reference`` symlink, a ``Control_Toolkit_ASF`` with working
``config_*.yml`` (the reference reads them relative to CWD at import time) and the CartPole cost plugins,
with the SI_Toolkit / tensorflow / watchdog shims on ``sys.path``.  Only usable where /root/reference exists
(this container); never on the GPU box."""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("CTK_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "Optimizers"))


def enter_workspace() -> str:
    """Create the workspace, chdir into it, extend sys.path.  Returns the workspace path."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE}")
    ws = tempfile.mkdtemp(prefix="ctk_ref_ws_")
    os.symlink(REFERENCE, os.path.join(ws, "Control_Toolkit"))
    shutil.copytree(os.path.join(HERE, "asf", "Control_Toolkit_ASF"), os.path.join(ws, "Control_Toolkit_ASF"))
    os.chdir(ws)
    for p in (REPO, HERE, ws):
        if p not in sys.path:
            sys.path.insert(0, p)
    return ws
