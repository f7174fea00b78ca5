"""Runs an UNMODIFIED reference script with rba_b200 plugged in:

    cd <reference checkout> && python -m rba_b200.compat.run evaluate_ood.py [the script's own arguments]

Equivalent to `python evaluate_ood.py ...` after `rba_b200.compat.plug_in()`; the script's directory is put first on
sys.path exactly as the interpreter would do for a script."""
import importlib.util
import os
import runpy
import sys
import types


def prefer_local_namespace_packages(script_dir):
    """The reference keeps helper code in directories without __init__.py (`datasets/`, imported as
    `from datasets.cityscapes import ...`, evaluate_ood.py:13-17).  Python resolves such namespace packages AFTER any
    regular package of the same name on sys.path (e.g. HuggingFace `datasets` in site-packages), whatever the path
    order.  Give the script's own directories precedence, as they have in the reference's environment."""
    done = []
    for name in sorted(os.listdir(script_dir)):
        d = os.path.join(script_dir, name)
        if not os.path.isdir(d) or not name.isidentifier() or os.path.exists(os.path.join(d, "__init__.py")):
            continue
        if not any(f.endswith(".py") for f in os.listdir(d)) or name in sys.modules:
            continue
        try:
            spec = importlib.util.find_spec(name)
        except (ImportError, ValueError):
            spec = None
        if spec is not None and spec.origin is not None and not os.path.abspath(spec.origin).startswith(d):
            m = types.ModuleType(name)
            m.__path__ = [d]
            sys.modules[name] = m
            done.append(name)
    return done


def run_script(script, args=()):
    script = os.path.abspath(script)
    sdir = os.path.dirname(script)
    sys.argv = [script] + list(args)
    if sdir not in sys.path[:1]:
        sys.path.insert(0, sdir)
    prefer_local_namespace_packages(sdir)
    runpy.run_path(script, run_name="__main__")


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m rba_b200.compat.run <script.py> [args...]")
    from rba_b200 import compat
    served = compat.plug_in()
    print(f"[rba_b200.compat] stand-ins serve: {', '.join(served) if served else '(none: all third parties are installed)'}",
          file=sys.stderr)
    run_script(argv[0], argv[1:])


if __name__ == "__main__":
    main()
