"""TEST-ONLY shim: the reference wraps hot functions in CompileTF / CompileAdaptive (tf.function /
torch.compile).  For golden-vector generation they are identities (eager, source-as-written semantics:
SURVEY.md section 8 'quirks' -- previous_input is read live)."""


def CompileTF(f):
    return f


def CompileAdaptive(f):
    return f
