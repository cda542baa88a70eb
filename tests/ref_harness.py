"""Import the reference (LeCAR-Lab/SPI-Active, /root/reference) in the build container and build its REAL env classes on
top of the `simulator=b200` plugin.  TEST INFRASTRUCTURE; nothing here is reachable from the product package, and nothing
that runs on the GPU box imports it (/root/reference does not exist there).

The reference's third-party imports that are not installable here (hydra, omegaconf, optuna, matplotlib, isaacgym, loguru,
rich, termcolor, easydict ...) are replaced by inert MagicMock modules; the three calls whose RESULT matters are given real
stand-ins: `hydra.compose` -> tests/mini_hydra.compose over [repo config/, reference config/], `OmegaConf.create` -> a plain
attribute dict, `hydra.utils.get_class` -> importlib.  Everything else that executes is the reference's unmodified code:
scripts/eval.py load_env_config / instantiate_env / evaluate_batch / apply_base_mass, BaseTask.__init__,
LeggedRobotBase._init_buffers / reset_all / step / _post_physics_step ...
"""
from __future__ import annotations

import importlib
import sys
import warnings
from pathlib import Path
from types import SimpleNamespace
from unittest import mock

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")

STUBBED = ["hydra", "hydra.core", "hydra.core.config_store", "hydra.utils", "hydra.core.hydra_config", "omegaconf",
           "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "mpl_toolkits",
           "mpl_toolkits.mplot3d", "termcolor", "easydict", "isaacgym", "optuna", "optuna.samplers", "loguru", "rich",
           "rich.progress", "ipdb", "pynput"]


def available() -> bool:
    return (REF / "scripts" / "eval.py").exists()


def get_class(path: str):
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


_R = None


def import_reference():
    global _R
    if _R is not None:
        return _R
    if not available():
        raise RuntimeError("/root/reference is not present")
    for p in (str(ROOT), str(ROOT / "tests"), str(REF), str(REF / "isaac_utils")):
        if p not in sys.path:
            sys.path.insert(0, p)
    stubbed_here = []
    for name in STUBBED:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = mock.MagicMock(name=name)
            stubbed_here.append(name)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    import scripts.eval as ev
    import scripts.mass_landscape as ml
    import scripts.mass_opt as mo
    import spigym.envs.base_task.base_task as bt
    from spigym.envs.legged_base_task.legged_robot_base import LeggedRobotBase
    from spigym.envs.locomotion.go2_omni import go2_omni_interface
    from spigym.envs.sysid.active_sysid_openloop import ActiveSysId_OpenLoop
    import spigym.utils.torch_utils as tu
    import mini_hydra
    search = [ROOT / "config", REF / "spigym" / "config"]
    if isinstance(ev.hydra, mock.MagicMock):
        ev.hydra.compose = lambda config_name, **kw: mini_hydra.compose(config_name, search)
    if isinstance(ev.OmegaConf, mock.MagicMock):
        ev.OmegaConf.create = mini_hydra._wrap
    if isinstance(bt.get_class, mock.MagicMock):
        bt.get_class = get_class
    # the reference modules keep their own references to the stubs; take them out of sys.modules again so that product code
    # probing for the real package (`import optuna` in landscape.optimize_mass) still sees it as absent
    for name in stubbed_here:
        sys.modules.pop(name, None)
    _R = SimpleNamespace(ev=ev, ml=ml, mo=mo, bt=bt, LeggedRobotBase=LeggedRobotBase, go2_omni=go2_omni_interface,
                         ActiveSysId=ActiveSysId_OpenLoop, tu=tu)
    return _R


def instantiate_env(num_envs: int, target: str, base_dir=Path("/tmp/spi_b200_ref_env"), seed: int = 0):
    """scripts/eval.py:313-321 instantiate_env, unmodified, with `simulator._target_` pointed at `target`
    (a BaseSimulator subclass path).  Returns the reference's LeggedRobotBase after set_is_evaluating() + reset_all()."""
    R = import_reference()
    import mini_hydra
    search = [ROOT / "config", REF / "spigym" / "config"]

    def compose(config_name, **kw):
        cfg = mini_hydra.compose(config_name, search)
        cfg.simulator["_target_"] = target
        return cfg
    R.ev.hydra.compose = compose
    return R.ev.instantiate_env(Path(base_dir), num_envs, "env/go2_test", seed)
