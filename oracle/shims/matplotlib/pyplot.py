"""No plotting in the oracle harness; any call is a no-op returning None."""


def __getattr__(name):
    def _noop(*a, **k):
        return None
    return _noop
