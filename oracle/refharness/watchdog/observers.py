"""TEST-ONLY empty stand-in for watchdog.observers (imported at reference Cost_Functions/CostFunctionUpdater.py:3)."""


class Observer:
    daemon = True

    def schedule(self, *a, **k):
        pass

    def start(self):
        pass

    def stop(self):
        pass

    def join(self, *a, **k):
        pass
