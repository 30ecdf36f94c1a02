"""TEST-ONLY empty stand-in for watchdog.events (imported at reference Cost_Functions/CostFunctionUpdater.py:4)."""


class FileSystemEventHandler:
    pass
