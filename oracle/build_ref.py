"""Recipe for oracle/_ref/: byte-compiles the UNMODIFIED reference modules where they lie under
/root/reference into marshalled code objects, <Module>.code (build outputs only -- no reference source
enters the repository; oracle/_ref/ is git-ignored but travels to the GPU box with the snapshot; the
files are not named .pyc because snapshot tools commonly drop those).

    python -m oracle.build_ref

TEST INFRASTRUCTURE ONLY: used by oracle/ref_harness.py (golden generation, the live
`requires_reference` cross-checks) and by bench.py's CPU legs to time the Python reference itself
on the host cores.  Rendering modules (pygame) are not compiled: they stay off the hot path.
"""
from __future__ import annotations

import marshal
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SOURCE_DIR = os.environ.get("SKILLSHOT_REFERENCE_DIR", "/root/reference")
MODULES = ("Projectile", "Player", "SkillshotGame", "SkillshotLearner", "InputHandler")
PLAYABLE_TICK = "playable_tick.marshal"       # the key-press -> move -> game_tick statements of skillshot_playable.py:51-64


def playable_tick_code(source_dir: str = None):
    """The two statements skillshot_playable.py runs per frame between event handling and drawing (lines 51-64: the loop
    that turns the InputHandler's key states into Player.move_* calls, then skillshotGame.game_tick()), cut out of the
    module's syntax tree and compiled on their own.  The script itself cannot be imported (it opens a pygame window at
    module scope); the statements refer to the globals `inputHandler` and `skillshotGame`."""
    import ast
    path = os.path.join(source_dir or SOURCE_DIR, "skillshot_playable.py")
    tree = ast.parse(open(path).read(), filename="skillshot_playable.py")
    loop = next(n for n in tree.body if isinstance(n, ast.While))
    keep = []
    for node in loop.body:
        text = ast.unparse(node)
        if isinstance(node, ast.For) and "inputHandler.get_inputs()" in text and "move_forwards" in text:
            keep.append(node)
        elif isinstance(node, ast.Expr) and text.strip() == "skillshotGame.game_tick()":
            keep.append(node)
    assert len(keep) == 2, "skillshot_playable.py no longer has the shape this recipe expects"
    return compile(ast.Module(body=keep, type_ignores=[]), "skillshot_playable.py", "exec")


def built() -> bool:
    return all(os.path.exists(os.path.join(OUT, m + ".code")) for m in MODULES) and os.path.exists(os.path.join(OUT, PLAYABLE_TICK))


def build(force: bool = False) -> bool:
    """True when oracle/_ref holds the compiled reference (built now or earlier)."""
    if not os.path.exists(os.path.join(SOURCE_DIR, "SkillshotGame.py")):
        return built()
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        src, dst = os.path.join(SOURCE_DIR, m + ".py"), os.path.join(OUT, m + ".code")
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")      # `is not 0` SyntaxWarning, SkillshotGame.py:44,54
                code = compile(open(src).read(), m + ".py", "exec")
            with open(dst, "wb") as f:
                marshal.dump(code, f)
    with open(os.path.join(OUT, PLAYABLE_TICK), "wb") as f:
        marshal.dump(playable_tick_code(), f)
    with open(os.path.join(OUT, "BUILT_FROM"), "w") as f:
        f.write("%s, python %s\n" % (SOURCE_DIR, sys.version.split()[0]))
    return True


def load_module(name: str):
    """Import reference module `name` from its marshalled code object (its own imports of sibling reference modules
    resolve through sys.modules, so load in the order of MODULES)."""
    import types
    if name in sys.modules:
        return sys.modules[name]
    with open(os.path.join(OUT, name + ".code"), "rb") as f:
        code = marshal.load(f)
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(OUT, name + ".code")
    sys.modules[name] = mod
    try:
        exec(code, mod.__dict__)
    except BaseException:
        del sys.modules[name]
        raise
    return mod


if __name__ == "__main__":
    print("oracle/_ref:", "ok" if build(force="--force" in sys.argv) else "reference not mounted")
