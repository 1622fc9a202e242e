"""The GPU-only tools, the examples and bench.py cannot run on the CPU test box; at least every one of them must
byte-compile and keep its imports resolvable."""
import ast
import importlib.util
import pathlib
import py_compile

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
SCRIPTS = sorted(list((ROOT / "tools").glob("*.py")) + list((ROOT / "examples").glob("*.py")) +
                 [ROOT / "bench.py", ROOT / "__graft_entry__.py"])


@pytest.mark.parametrize("path", SCRIPTS, ids=lambda p: p.name)
def test_script_compiles_and_imports_resolve(path):
    py_compile.compile(str(path), doraise=True)
    tree = ast.parse(path.read_text())
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("pipegoose_b200"):
            spec = importlib.util.find_spec(node.module)
            assert spec is not None, f"{path.name}: cannot resolve {node.module}"
            mod = importlib.import_module(node.module)
            for alias in node.names:
                assert alias.name == "*" or hasattr(mod, alias.name) or \
                    importlib.util.find_spec(f"{node.module}.{alias.name}") is not None, \
                    f"{path.name}: {node.module} has no attribute {alias.name}"
