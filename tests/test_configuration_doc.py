"""docs/CONFIGURATION.md lists every ``PIPEGOOSE_B200_*`` environment switch the code reads, and nothing the code does
not read (a switch that is set by a tool but read nowhere — PIPEGOOSE_B200_FUSED_MOE was one — silently does nothing)."""
import pathlib
import re

ROOT = pathlib.Path(__file__).resolve().parent.parent
NAME = re.compile(r"PIPEGOOSE_B200_[A-Z0-9_]+")
READ = re.compile(r"""(?:environ\.get\(|environ\[|getenv\()\s*["'](PIPEGOOSE_B200_[A-Z0-9_]+)["']""")


def _sources():
    files = [ROOT / "bench.py"]
    for pattern in ("*.py", "*.cu", "*.cuh", "*.cpp", "*.h"):
        files += [f for f in (ROOT / "pipegoose_b200").rglob(pattern) if "build" not in f.parent.name]
    return files


def test_every_switch_that_is_read_is_documented_and_vice_versa():
    read = set()
    for f in _sources():
        read |= set(READ.findall(f.read_text(errors="ignore")))
    documented = set(NAME.findall((ROOT / "docs" / "CONFIGURATION.md").read_text()))
    assert read - documented == set(), f"read by the code, missing in docs/CONFIGURATION.md: {sorted(read - documented)}"
    assert documented - read == set(), f"documented, but read nowhere: {sorted(documented - read)}"


def test_switches_that_tools_set_are_read_somewhere():
    read = set()
    for f in _sources():
        read |= set(READ.findall(f.read_text(errors="ignore")))
    mentioned = set()
    for f in [ROOT / "bench.py"] + sorted((ROOT / "tools").glob("*")) + sorted((ROOT / "examples").glob("*.py")):
        if f.is_file():
            mentioned |= set(NAME.findall(f.read_text(errors="ignore")))
    assert mentioned - read == set(), f"set or mentioned by a tool, read nowhere: {sorted(mentioned - read)}"
