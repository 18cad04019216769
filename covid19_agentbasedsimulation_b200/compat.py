"""Drop-in aliasing: makes `import covid_abs...` resolve to this package, so that scripts written against the
reference (`from covid_abs.abs import Simulation`, `from covid_abs.experiments import batch_experiment`,
`from covid_abs.network.graph_abs import GraphSimulation`, ...) run unmodified on the B200 path.

    import covid19_agentbasedsimulation_b200.compat as compat
    compat.install()                      # before the script's own `import covid_abs...`
    ... user script ...
    compat.uninstall()                    # restores whatever `covid_abs` modules were there before

or run an unmodified script through it:

    python -m covid19_agentbasedsimulation_b200.compat user_script.py [args...]

Module map (reference layout, SURVEY.md section 1 -> here):
    covid_abs.abs                 -> covid19_agentbasedsimulation_b200.abs
    covid_abs.agents              -> .agents
    covid_abs.common              -> .common
    covid_abs.experiments         -> .experiments          (batch_experiment; the plotting helpers are out of scope)
    covid_abs.network             -> .network
    covid_abs.network.graph_abs   -> .network.graph_abs
    covid_abs.network.agents      -> .network.agents
    covid_abs.network.util        -> .network.util
`covid_abs.graphics` is presentation code (matplotlib animations) and is not provided: importing it raises
ImportError with that explanation instead of silently falling back to the CPU reference.
"""
import importlib
import importlib.abc
import importlib.machinery
import runpy
import sys

_PKG = "covid19_agentbasedsimulation_b200"
_MAP = {
    "covid_abs": _PKG,
    "covid_abs.abs": _PKG + ".abs",
    "covid_abs.agents": _PKG + ".agents",
    "covid_abs.common": _PKG + ".common",
    "covid_abs.experiments": _PKG + ".experiments",
    "covid_abs.network": _PKG + ".network",
    "covid_abs.network.graph_abs": _PKG + ".network.graph_abs",
    "covid_abs.network.agents": _PKG + ".network.agents",
    "covid_abs.network.util": _PKG + ".network.util",
}
_saved = None


class _Blocker(importlib.abc.MetaPathFinder):
    """Any other `covid_abs.*` import fails loudly while the alias is installed (no silent CPU path)."""

    def find_spec(self, fullname, path, target=None):
        if fullname == "covid_abs" or fullname.startswith("covid_abs."):
            if fullname not in _MAP:
                raise ImportError("{} is not part of the B200 drop-in (covid19_agentbasedsimulation_b200.compat): only "
                                  "the simulation step, its statistics and batch_experiment are provided".format(fullname))
        return None


_blocker = _Blocker()


def install():
    """Alias the `covid_abs` module tree to this package.  Idempotent."""
    global _saved
    if _saved is not None:
        return
    _saved = {name: mod for name, mod in sys.modules.items() if name == "covid_abs" or name.startswith("covid_abs.")}
    for name in list(_saved):
        del sys.modules[name]
    for alias, real in _MAP.items():
        sys.modules[alias] = importlib.import_module(real)
    sys.meta_path.insert(0, _blocker)


def uninstall():
    """Undo install(): the previously imported `covid_abs` modules (if any) come back."""
    global _saved
    if _saved is None:
        return
    for alias in _MAP:
        sys.modules.pop(alias, None)
    if _blocker in sys.meta_path:
        sys.meta_path.remove(_blocker)
    sys.modules.update(_saved)
    _saved = None


def installed():
    return _saved is not None


class aliased(object):
    """Context manager form: `with compat.aliased(): exec(user_code)`."""

    def __enter__(self):
        install()
        return self

    def __exit__(self, *exc):
        uninstall()
        return False


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    install()
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
