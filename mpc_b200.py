"""Import shim: the package directory is named ``carnd-mpc-project_b200`` (not a Python identifier),
so it is loaded here under the module name ``carnd_mpc_project_b200`` and re-exported.

    import mpc_b200 as mpc          # mpc.Solver, mpc.config_from_json_text, mpc.workloads ...
"""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "carnd-mpc-project_b200")
_NAME = "carnd_mpc_project_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

_pkg = sys.modules[_NAME]
from carnd_mpc_project_b200 import *  # noqa: F401,F403,E402
from carnd_mpc_project_b200 import (MpcConfig, MpcError, Solver, MultiSolver, build, lib, LIB_PATH, EXPORTS,  # noqa: E402,F401
                                    config_defaults, config_from_json_file, config_from_json_text, config_from_cli,
                                    STATUS_NAMES, STATUS_SUCCESS, KERNEL_AUTO, KERNEL_LANE, KERNEL_COOP,
                                    LANE_MIN_BATCH, MpcRunAux, MpcTelemetry, telemetry_parse, run_prepare, run_finish, compute_throttle,
                                    vehicle_move)
import carnd_mpc_project_b200.workloads as workloads  # noqa: E402,F401
import carnd_mpc_project_b200.sharding as sharding  # noqa: E402,F401
