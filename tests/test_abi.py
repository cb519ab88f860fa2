"""The C-ABI library loads and exports every symbol include/jutul_b200.h declares; the ctypes
prototype table matches the header; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "jutul_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(J):
    lib = J._lib.load()
    names = _header_functions()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), f"libjutul_b200.so does not export {n}"


def test_prototype_table_matches_header(J):
    assert sorted(J._lib.PROTOTYPES) == _header_functions()
    assert J._lib.load().jb_version() == 100


def _header_prototypes():
    """name -> (return kind, [argument kinds]) parsed from the header; kinds: p (any pointer), i32, i64, f64, void."""
    src = open(os.path.join(ROOT, "include", "jutul_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def kind(t):
        t = t.strip()
        if "*" in t:
            return "p"
        base = re.sub(r"\bconst\b", "", t).split()[0]
        return {"int32_t": "i32", "int64_t": "i64", "double": "f64", "void": "void"}[base]
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(jb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        out[name] = (kind(ret), [] if args in ("", "void") else [kind(a) for a in args.split(",")])
    return out


def test_ctypes_prototypes_match_header_signatures(J):
    """Argument by argument: pointer / int32 / int64 / double in the ctypes table = the C declaration."""
    protos = _header_prototypes()
    assert sorted(protos) == _header_functions()

    def ck(t):
        return "void" if t is None else {C.c_int32: "i32", C.c_int64: "i64", C.c_double: "f64"}.get(t, "p")
    for name, (ret, args) in J._lib.PROTOTYPES.items():
        assert (ck(ret), [ck(a) for a in args]) == protos[name], name


def test_julia_ccall_signatures_match_header():
    """Every ccall of julia/JutulB200.jl passes the return type and the argument types (Ptr / Ref / Cstring, Int32, Int64,
    Float64) the C declaration has, in the same order — the check a first run under Julia would otherwise make."""
    protos = _header_prototypes()
    jl = open(os.path.join(ROOT, "jutul.jl_b200", "julia", "JutulB200.jl")).read()

    def jk(t):
        t = t.strip()
        if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
            return "p"
        return {"Int32": "i32", "Cint": "i32", "Int64": "i64", "Float64": "f64", "Cdouble": "f64", "Cvoid": "void"}[t]
    n = 0
    for m in re.finditer(r"ccall\(\(:(jb_[a-z0-9_]+), LIB\),\s*([\w{}]+),\s*\((.*?)\)\s*[,)]", jl, flags=re.S):
        name, ret, tup = m.groups()
        assert (jk(ret), [jk(t) for t in tup.split(",") if t.strip()]) == protos[name], name
        n += 1
    assert n == jl.count("ccall(") and n >= 70


def test_julia_glue_blocks_and_brackets_balance():
    """Structural lint of the unexecuted glue: with strings and comments removed, every bracket closes and every block opener
    (function / if / for / while / let / struct / begin / try / do / module / macro / quote) has its `end`. Generators'
    for / if inside brackets and `a[end]` are not blocks."""
    src = open(os.path.join(ROOT, "jutul.jl_b200", "julia", "JutulB200.jl")).read()
    src = re.sub(r'"""(?:.|\n)*?"""', '""', src)
    src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)
    src = re.sub(r"#=(?:.|\n)*?=#", "", src)
    src = re.sub(r"#[^\n]*", "", src)
    pairs, stack = {")": "(", "]": "[", "}": "{"}, []
    for ch in src:
        if ch in "([{":
            stack.append(ch)
        elif ch in ")]}":
            assert stack and stack.pop() == pairs[ch]
    assert not stack
    openers = {"function", "if", "for", "while", "let", "struct", "begin", "try", "do", "module", "macro", "quote"}
    dp = db = 0
    blocks, prev = [], None
    for t in re.findall(r"[A-Za-z_][A-Za-z_0-9!]*|\S", src):
        if t == "(":
            dp += 1
        elif t == ")":
            dp -= 1
        elif t == "[":
            db += 1
        elif t == "]":
            db -= 1
        elif t in openers and prev not in (":", "."):
            if not (t in ("for", "if") and (dp > 0 or db > 0)):
                blocks.append(db)
        elif t == "end" and prev != ":":
            if not (db > 0 and (not blocks or blocks[-1] < db)):
                assert blocks, "unmatched end"
                blocks.pop()
        prev = t
    assert not blocks


def test_sm100a_cubin_present(J):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", J._lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_cpu_fallback(J):
    with pytest.raises(J.JutulB200Error) as e:
        J.B200Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def build_c_harness():
    """gcc tests/abi_c_harness.c -> tests/_build/abi_c_harness (plain C, dlopen; stand-in for Julia's ccall)."""
    import subprocess
    out = os.path.join(ROOT, "tests", "_build", "abi_c_harness")
    src = os.path.join(ROOT, "tests", "abi_c_harness.c")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-o", out, src, "-ldl"])
    return out


def test_c_harness_resolves_every_entry_point(J):
    """The header compiles as C and every entry point the reference-side glue binds for a Newton iteration resolves through
    dlopen/dlsym, without Python in the process."""
    import subprocess
    exe = build_c_harness()
    res = subprocess.run([exe, J._lib.SO_PATH, "symbols"], capture_output=True, text=True)
    assert res.returncode == 0 and "ok version=100" in res.stdout, res.stderr


def test_julia_glue_binds_only_exported_symbols(J):
    """Every ccall in julia/JutulB200.jl names a function the header declares (the glue cannot be executed here)."""
    src = open(os.path.join(ROOT, "jutul.jl_b200", "julia", "JutulB200.jl")).read()
    called = set(re.findall(r"ccall\(\(:(jb_[a-z0-9_]+), LIB\)", src))
    assert len(called) >= 35
    assert called <= set(_header_functions()), called - set(_header_functions())
    assert 'error("device-resident path' not in src                      # no stub methods
    for method in ("function linear_solve!(sys::LSystem, krylov::GenericKrylov, context::B200Context",
                   "function update_linearized_system_equation!(nz, r, model::B200Model",
                   "function update_equation!(s::B200TPFAStorage", "function convergence_criterion(model::B200Model",
                   "function update_primary_variable!(state, p::Jutul.ScalarVariable", "function apply!(x::Vector{Float64}, F::B200Factor",
                   "function update_preconditioner!(ilu::ILUZeroPreconditioner, A::B200Matrix"):
        assert method in src, method


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "jutul.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "oracle.cpp" not in txt and "liborc" not in txt, f
