"""Import the REAL reference (ultralytics fork at /root/reference) in this container.

Two shims are needed, neither edits the reference tree:
  1. a `matplotlib` stub on sys.path (the reference hard-imports it);
  2. `RepVGGBlock` registered in `ultralytics.nn.tasks` and in `parse_model`'s `base_modules` at runtime.
The reference does not exist on the GPU box: only tools/make_golden.py and the `reference`-marked tests use this.
"""
import inspect
import os
import sys
import tempfile
from pathlib import Path

REF_ROOT = Path(os.environ.get("DRONEYOLO_REF", "/root/reference"))


def available() -> bool:
    return (REF_ROOT / "ultralytics" / "nn" / "tasks.py").is_file()


def load():
    """Returns the patched `ultralytics.nn.tasks` module of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp(prefix="ulcfg"))
    os.environ.setdefault("YOLO_OFFLINE", "1")
    sys.dont_write_bytecode = True
    stubs = str(Path(__file__).resolve().parent / "stubs")
    for p in (str(REF_ROOT), stubs):
        if p not in sys.path:
            sys.path.insert(0, p)
    from ultralytics.nn import tasks
    from ultralytics.nn.modules.block import RepVGGBlock

    if not getattr(tasks, "_droneyolo_shim", False):
        tasks.RepVGGBlock = RepVGGBlock                        # name lookup, tasks.py:1017
        src = inspect.getsource(tasks.parse_model)
        marker = "            Classify,\n"                     # first member of base_modules, tasks.py:956
        assert marker in src
        src = src.replace(marker, marker + "            RepVGGBlock,\n", 1)
        exec(compile(src, tasks.__file__, "exec"), tasks.__dict__)
        tasks._droneyolo_shim = True
    return tasks
