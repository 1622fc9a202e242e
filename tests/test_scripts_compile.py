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


def _undefined_names(path):
    """Names that a function reads as globals although the module never binds them (what pyflakes calls an
    undefined name) — the bug class that only shows up when the GPU-only code path finally runs."""
    import builtins
    import symtable

    src = path.read_text()
    top = symtable.symtable(src, str(path), "exec")
    if any(isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names) for n in ast.walk(ast.parse(src))):
        return []
    bound = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    bound |= set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__builtins__", "__spec__", "__path__", "__class__"}
    # names bound through ``global x`` inside a function
    stack, missing = [top], []
    tables = []
    while stack:
        t = stack.pop()
        tables.append(t)
        stack.extend(t.get_children())
    for t in tables:
        for s in t.get_symbols():
            if t is not top and s.is_declared_global() and s.is_assigned():
                bound.add(s.get_name())
    for t in tables:
        if t is top:
            continue
        for s in t.get_symbols():
            if s.is_global() and s.is_referenced() and s.get_name() not in bound:
                missing.append(f"{t.get_name()}:{s.get_name()}")
    for s in top.get_symbols():
        if s.is_referenced() and not (s.is_assigned() or s.is_imported() or s.is_namespace()) and s.get_name() not in bound:
            missing.append(f"<module>:{s.get_name()}")
    return missing


PACKAGE = sorted((ROOT / "pipegoose_b200").rglob("*.py"))


@pytest.mark.parametrize("path", SCRIPTS + PACKAGE, ids=lambda p: str(p.relative_to(ROOT)))
def test_no_undefined_names(path):
    assert _undefined_names(path) == []


@pytest.mark.parametrize("path", SCRIPTS + PACKAGE, ids=lambda p: str(p.relative_to(ROOT)))
def test_module_alias_attributes_and_native_bindings_exist(path):
    """``K.gemm_tn`` / ``S.SymmetricWorkspace`` / ``native().attention_fwd`` style references: the attribute exists in
    the aliased ``pipegoose_b200`` module, and every ``native().<name>`` is a function the extension binds."""
    import re

    bound = set(re.findall(r'm\.def\("([a-z_0-9]+)"', (ROOT / "pipegoose_b200" / "csrc" / "bindings.cpp").read_text()))
    tree = ast.parse(path.read_text())
    aliases = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            aliases.update({a.asname: a.name for a in node.names if a.asname and a.name.startswith("pipegoose_b200")})
        elif isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("pipegoose_b200"):
            for a in node.names:
                if a.name != "*" and importlib.util.find_spec(node.module) is not None:
                    try:
                        is_module = importlib.util.find_spec(f"{node.module}.{a.name}") is not None
                    except ModuleNotFoundError:
                        is_module = False
                    if is_module:
                        aliases[a.asname or a.name] = f"{node.module}.{a.name}"
    missing = []
    for node in ast.walk(tree):
        if not isinstance(node, ast.Attribute):
            continue
        if isinstance(node.value, ast.Name) and node.value.id in aliases:
            if not hasattr(importlib.import_module(aliases[node.value.id]), node.attr):
                missing.append(f"line {node.lineno}: {node.value.id}.{node.attr}")
        elif isinstance(node.value, ast.Call) and isinstance(node.value.func, ast.Name) and node.value.func.id == "native":
            if node.attr not in bound:
                missing.append(f"line {node.lineno}: native().{node.attr}")
    assert missing == []


def test_native_call_sites_pass_a_valid_number_of_arguments():
    """Every ``native().<fn>(...)`` call gives the bound C++ function between its required and its total number of
    arguments (parsed from csrc/bindings.cpp) — the arity errors that otherwise only surface on a GPU box."""
    import re

    src = (ROOT / "pipegoose_b200" / "csrc" / "bindings.cpp").read_text()
    n_params = {}
    for m in re.finditer(r"^[\w:<>\s\*&]+?\b(\w+)\(([^{;]*?)\)\s*\{", src, re.M):
        depth, n = 0, (1 if m.group(2).strip() else 0)
        for ch in m.group(2):
            depth += ch in "<([" 
            depth -= ch in ">)]"
            n += ch == "," and depth == 0
        n_params[m.group(1)] = n
    bounds = {}
    for m in re.finditer(r'm\.def\("(\w+)",\s*&(\w+)([^;]*);', src):
        total = n_params.get(m.group(2))
        if total is not None:
            bounds[m.group(1)] = (total - len(re.findall(r'py::arg\("\w+"\)\s*=', m.group(3))), total)
    assert len(bounds) >= 20, "bindings.cpp was not parsed"
    wrong = []
    for path in SCRIPTS + PACKAGE:
        for node in ast.walk(ast.parse(path.read_text())):
            if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Call)
                    and isinstance(node.func.value.func, ast.Name) and node.func.value.func.id == "native"):
                continue
            if node.func.attr not in bounds or any(isinstance(a, ast.Starred) for a in node.args):
                continue
            given = len(node.args) + len(node.keywords)
            lo, hi = bounds[node.func.attr]
            if not lo <= given <= hi:
                wrong.append(f"{path.relative_to(ROOT)}:{node.lineno} {node.func.attr}: {given} args, takes {lo}..{hi}")
    assert wrong == []


@pytest.mark.parametrize("path", SCRIPTS + PACKAGE, ids=lambda p: str(p.relative_to(ROOT)))
def test_calls_into_aliased_modules_bind_to_the_callee_signature(path):
    """``K.layernorm_fwd(x, g, b, eps, out=stage)`` / ``PF.embedding_positions(...)`` style calls: the positional count and
    every keyword are accepted by the function's signature (checked with ``inspect.signature(...).bind``), including the
    calls that only execute on a GPU box."""
    import inspect

    tree = ast.parse(path.read_text())
    aliases = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            aliases.update({a.asname: a.name for a in node.names if a.asname and a.name.startswith("pipegoose_b200")})
        elif isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("pipegoose_b200"):
            for a in node.names:
                try:
                    if a.name != "*" and importlib.util.find_spec(f"{node.module}.{a.name}") is not None:
                        aliases[a.asname or a.name] = f"{node.module}.{a.name}"
                except ModuleNotFoundError:
                    pass
    # names imported directly: ``from pipegoose_b200.x import f`` (module-level imports only: no shadowing games)
    direct = {}
    for node in tree.body:
        if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("pipegoose_b200"):
            for a in node.names:
                if a.name != "*" and (a.asname or a.name) not in aliases:
                    direct[a.asname or a.name] = (node.module, a.name)
    assigned = {t.id for n in ast.walk(tree) if isinstance(n, (ast.Assign, ast.AugAssign, ast.AnnAssign))
                for t in ast.walk(n.targets[0] if isinstance(n, ast.Assign) else n.target) if isinstance(t, ast.Name)}
    assigned |= {a.arg for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.Lambda)) for a in n.args.args + n.args.kwonlyargs}
    wrong = []
    for node in ast.walk(tree):
        if not isinstance(node, ast.Call):
            continue
        if isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name) and node.func.value.id in aliases:
            target = getattr(importlib.import_module(aliases[node.func.value.id]), node.func.attr, None)
            label = f"{node.func.value.id}.{node.func.attr}"
        elif isinstance(node.func, ast.Name) and node.func.id in direct and node.func.id not in assigned:
            module, name = direct[node.func.id]
            target = getattr(importlib.import_module(module), name, None)
            label = node.func.id
        else:
            continue
        if any(isinstance(a, ast.Starred) for a in node.args) or any(k.arg is None for k in node.keywords):
            continue
        if not (inspect.isfunction(target) or inspect.isclass(target)):
            continue
        try:
            sig = inspect.signature(target)
            sig.bind(*[None] * len(node.args), **{k.arg: None for k in node.keywords})
        except TypeError as e:
            wrong.append(f"line {node.lineno}: {label}: {e}")
        except ValueError:
            pass  # no signature available (builtins)
    assert wrong == []
