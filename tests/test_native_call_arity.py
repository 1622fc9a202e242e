"""Every Python call of an extension function passes a number of arguments (and keyword names) the C++ signature accepts.
The extension only loads on a GPU box, so a call site on a rarely taken path (an opt-in switch, an error path) with one
argument too many would otherwise be found by the first user who takes it."""
import ast
import pathlib
import re

ROOT = pathlib.Path(__file__).resolve().parent.parent / "pipegoose_b200"


def _cxx_sources() -> str:
    files = sorted((ROOT / "csrc").glob("*.c*")) + sorted((ROOT / "csrc").glob("*.h"))
    return "\n".join(f.read_text() for f in files if f.suffix in (".cu", ".cuh", ".cpp", ".h"))


def _split_params(body: str):
    if not body.strip():
        return []
    out, depth, cur = [], 0, ""
    for ch in body:
        depth += ch in "(<[{"
        depth -= ch in ")>]}"
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    return out + [cur]


def _signature(src: str, fn: str):
    """Parameters of the C++ function ``fn`` (its definition or declaration)."""
    for m in re.finditer(r"[\w:<>\*&\s]+?\b%s\s*\(" % re.escape(fn), src):
        i = j = m.end()
        depth = 1
        while depth and j < len(src):
            depth += src[j] == "("
            depth -= src[j] == ")"
            j += 1
        k = j
        while k < len(src) and src[k] in " \n\t":
            k += 1
        if k >= len(src) or src[k] not in "{;":
            continue
        params = _split_params(src[i:j - 1])
        if all(re.search(r"[\w>\*&]\s+\w+\s*(=.*)?$", p.strip()) for p in params):
            return params
    return None


def _bound_signatures():
    src = _cxx_sources()
    bind = (ROOT / "csrc" / "bindings.cpp").read_text()
    sigs = {}
    for name, fn in re.findall(r'm\.def\("([a-z_0-9]+)",\s*&([a-zA-Z_0-9]+)', bind):
        params = _signature(src, fn)
        assert params is not None, f"cannot find the C++ signature of {fn} (bound as {name})"
        names = [re.search(r"(\w+)\s*(=.*)?$", p.strip()).group(1) for p in params]
        sigs[name] = (len(params) - sum("=" in p for p in params), len(params), names)
    for m in re.finditer(r'm\.def\("([a-z_0-9]+)",\s*&\w+,(.*?)\);', bind, re.S):     # py::arg lists carry the defaults
        args = re.findall(r'py::arg\("(\w+)"\)(\s*=\s*[^,)]+)?', m.group(2))
        if args:
            sigs[m.group(1)] = (len(args) - sum(bool(d) for _, d in args), len(args), [a for a, _ in args])
    return sigs


def _is_extension_object(v) -> bool:
    if isinstance(v, ast.Call):
        return getattr(v.func, "id", getattr(v.func, "attr", "")) in ("native", "_C")
    if isinstance(v, ast.Name):
        return v.id in ("_n", "n", "mod", "ext", "nat")
    return isinstance(v, ast.Attribute) and v.attr in ("_n", "_native")


def test_extension_call_sites_match_the_cxx_signatures():
    sigs = _bound_signatures()
    assert len(sigs) >= 30
    checked, wrong = 0, []
    for path in sorted(ROOT.rglob("*.py")):
        for node in ast.walk(ast.parse(path.read_text())):
            if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in sigs
                    and _is_extension_object(node.func.value)):
                continue
            if any(isinstance(a, ast.Starred) for a in node.args) or any(k.arg is None for k in node.keywords):
                continue            # *args / **kwargs: built at run time (ops/kernels.py's gemm wrappers)
            lo, hi, names = sigs[node.func.attr]
            n_pos, kws = len(node.args), [k.arg for k in node.keywords]
            checked += 1
            if not (lo <= n_pos + len(kws) <= hi and all(k in names[n_pos:] for k in kws)):
                wrong.append(f"{path.relative_to(ROOT)}:{node.lineno} {node.func.attr}({n_pos} positional, {kws}) "
                             f"vs C++ {lo}..{hi} {names}")
    assert checked >= 40, f"the audit found only {checked} call sites: its idea of an extension object is out of date"
    assert wrong == [], "\n".join(wrong)
