"""Mechanical Python-2 -> Python-3 conversion of the reference's control loop, done in a TEMP DIR at test time
(build container only — nothing of the reference is copied into the repository).

/root/reference/src/agent.py and statistics.py are the CALLERS of the hot path; they are
Python 2 (`xrange`, `<>`, `time.clock`, `sys.maxint`, binary-mode csv) and there is no python2 here.  The five
token substitutions below are the whole conversion; `convert()` verifies that nothing else differs from the
original text, line for line.  INTEGRATION.md §1 points maintainers at this recipe."""
import importlib.util
import os
import re
import sys
import types

REFERENCE_SRC = "/root/reference/src"

# (pattern, replacement) — every one is a spelling change with identical semantics on the paths the loop takes
SUBSTITUTIONS = [
    (re.compile(r"\bxrange\("), "range("),                         # agent.py:32,91,100,110,122,129
    (re.compile(r"<>"), "!="),                                     # agent.py:70
    (re.compile(r"\btime\.clock\(\)"), "time.process_time()"),     # statistics.py:43,47,74 (CPU time on Linux/py2)
    (re.compile(r"\bsys\.maxint\b"), "sys.maxsize"),               # statistics.py:52-53
    (re.compile(r'open\(self\.csv_name, "wb"\)'), 'open(self.csv_name, "w", newline="")'),   # statistics.py:21
]


def convert(dst_dir, names=("agent.py", "statistics.py")):
    """Write converted copies of `names` into dst_dir; returns {name: number of changed lines}."""
    changed = {}
    for name in names:
        with open(os.path.join(REFERENCE_SRC, name)) as f:
            src = f.read()
        out = src
        for pat, rep in SUBSTITUTIONS:
            out = pat.sub(rep, out)
        a, b = src.splitlines(), out.splitlines()
        assert len(a) == len(b), "the conversion must not add or remove lines"
        diff = [(x, y) for x, y in zip(a, b) if x != y]
        for x, y in diff:                       # every changed line is explained by the substitutions alone
            z = x
            for pat, rep in SUBSTITUTIONS:
                z = pat.sub(rep, z)
            assert z == y
        compile(out, name, "exec")              # and the result is valid Python 3
        with open(os.path.join(dst_dir, name), "w") as f:
            f.write(out)
        changed[name] = len(diff)
    return changed


def load_module(path, name, injected=None):
    """Import a converted file under a private module name; `injected` = {module name: module object} made
    importable while it executes (agent.py does `from state_buffer import StateBuffer`)."""
    saved = {}
    for k, v in (injected or {}).items():
        saved[k] = sys.modules.get(k)
        sys.modules[k] = v
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load_agent_and_statistics(dst_dir, state_buffer_class, tag="ref"):
    """-> (Agent, Statistics) classes of the converted reference, wired to the given StateBuffer class."""
    convert(dst_dir)
    shim = types.ModuleType("state_buffer")
    shim.StateBuffer = state_buffer_class
    agent = load_module(os.path.join(dst_dir, "agent.py"), "_refconv_agent_" + tag, {"state_buffer": shim})
    stats = load_module(os.path.join(dst_dir, "statistics.py"), "_refconv_statistics_" + tag)
    return agent.Agent, stats.Statistics


def load_reference_replay_and_statebuffer():
    """The UNMODIFIED reference replay_memory.py / state_buffer.py under the numpy shim of SURVEY §8(c)."""
    import numpy as np
    rm = load_module(os.path.join(REFERENCE_SRC, "replay_memory.py"), "_ref_replay_memory")
    ns = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    ns.integer = np.int64            # what the abstract np.integer dtype meant on the reference's numpy
    rm.np = ns
    sb = load_module(os.path.join(REFERENCE_SRC, "state_buffer.py"), "_ref_state_buffer")
    return rm.ReplayMemory, sb.StateBuffer
