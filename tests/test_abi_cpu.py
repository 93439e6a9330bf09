"""CPU-side checks of the C ABI: the library builds/loads, exports every symbol include/consolver.h declares,
the ctypes signatures cover them, and argument validation answers with error codes before touching CUDA."""
import os
import re
import subprocess

import pytest

from consolver_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "consolver.h")).read()
    return sorted(set(re.findall(r"CONSOLVER_API\s+[\w\s\*]+?\b(consolver_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for n in ("consolver_policy_f32", "consolver_policy_table_f32", "consolver_policy_sample_f32", "consolver_step_sd", "consolver_step_fm", "consolver_sd_policy_and_step",
              "consolver_abi_version", "consolver_error_string"):
        assert n in names


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(consolver_\w+)", out))
    for n in _declared():
        assert n in exported, f"{n} declared in consolver.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
        assert hasattr(lib, n)
    hdr = open(os.path.join(ROOT, "include", "consolver.h")).read()
    assert lib.consolver_abi_version() == int(re.search(r"#define CONSOLVER_ABI_VERSION (\d+)", hdr).group(1))
    assert lib.consolver_abi_version() == _lib.ABI_VERSION


def test_library_carries_the_hash_of_the_header_it_was_built_from():
    from consolver_b200 import build

    lib = _lib.load()
    assert lib.consolver_abi_hash() == build.header_hash()
    assert build.is_current(), "libconsolver.so.stamp does not match the sources"
    # the hash the loader computes is FNV-1a 64 of the header bytes
    assert build.fnv1a64(b"") == 0xCBF29CE484222325 and build.fnv1a64(b"a") == 0xAF63DC4C8601EC8C


def test_a_library_built_from_an_older_header_is_rejected(tmp_path):
    """ABI hygiene: a stale .so must not be bound.  Three stand-ins for 'older': (1) a library whose embedded header hash
    differs, (2) one that predates consolver_abi_hash altogether, (3) one with another ABI version."""
    def make(name, body):
        src = tmp_path / (name + ".c")
        src.write_text(body)
        so = tmp_path / (name + ".so")
        subprocess.run(["gcc", "-shared", "-fPIC", "-o", str(so), str(src)], check=True)
        return str(so)

    v = _lib.ABI_VERSION
    other_hash = make("other_hash", f"int consolver_abi_version(void){{return {v};}}\n"
                                    "unsigned long long consolver_abi_hash(void){return 0x1234ull;}\n")
    with pytest.raises(_lib.ConsolverError, match="different include/consolver.h"):
        _lib.bind(other_hash)
    no_hash = make("no_hash", "int consolver_abi_version(void){return 1;}\n")
    with pytest.raises(_lib.ConsolverError, match="predates"):
        _lib.bind(no_hash)
    old_version = make("old_version", f"int consolver_abi_version(void){{return {v - 1};}}\n"
                                      "unsigned long long consolver_abi_hash(void){return 0;}\n")
    with pytest.raises(_lib.ConsolverError, match="ABI version"):
        _lib.bind(old_version)


def test_a_failed_rebuild_never_falls_back_to_the_existing_library(monkeypatch):
    """_lib.load() with a stale stamp and no working compiler raises instead of binding the old .so"""
    from consolver_b200 import build

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(build, "is_current", lambda: False)
    monkeypatch.setattr(build, "_nvcc", lambda: (_ for _ in ()).throw(RuntimeError("nvcc not found")))
    monkeypatch.delenv("CONSOLVER_NO_AUTOBUILD", raising=False)
    with pytest.raises(_lib.ConsolverError, match="does not fall back to a stale library"):
        _lib.load()


def test_header_compiles_as_plain_c():
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "consolver.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_is_built_for_sm_100a():
    r = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout


def test_argument_validation_returns_error_codes_without_a_gpu():
    lib = _lib.load()
    # null weights
    rc = lib.consolver_policy_f32(*([None] * 7), 0.0, 0.0, 999.0, 1.0, None, 0, None, None, 1, 256, 3, 11, 4, 0, 1, 0,
                                  *([None] * 7), None)
    assert rc == -1
    rc = lib.consolver_step_sd(0, None, None, 0.0, None, None, 1, None, None, None, 0, None, 6, 4, 1.0, 0.0, 1.0, 0.0,
                               0, 1, 16, None)
    assert rc == -1
    one = 16  # any non-null fake address: validation must fail on sizes before any dereference / launch
    rc = lib.consolver_step_sd(0, one, None, 0.0, None, None, 5, one, one, None, 0, one, 6, 4, 1.0, 0.0, 1.0, 0.0, 0, 1,
                               16, None)
    assert rc == -2  # n_hist > order_dim
    rc = lib.consolver_step_sd(0, one, None, 0.0, None, None, 1, one, one, None, 0, one, 6, 9, 1.0, 0.0, 1.0, 0.0, 0, 1,
                               16, None)
    assert rc == -2  # order_dim > CONSOLVER_MAX_ORDER
    rc = lib.consolver_step_sd(7, one, None, 0.0, None, None, 1, one, one, None, 0, one, 6, 4, 1.0, 0.0, 1.0, 0.0, 0, 1,
                               16, None)
    assert rc == -4  # dtype
    rc = lib.consolver_step_fm(2, 1, one, None, None, 1, one, one, None, 0, one, 6, 4, -0.1, 0, 1, 16, None)
    assert rc == -4  # x_dtype must be dtype or f32
    rc = lib.consolver_policy_f32(*([one] * 7), 0.0, 0.0, 999.0, 1.0, None, 0, one, one, 1, 256, 3, 11, 4, 0, 1, 0,
                                  *([one] * 7), None)
    assert rc == -1  # both q and idx_in
    rc = lib.consolver_policy_f32(*([one] * 7), 0.0, 0.0, 999.0, 1.0, None, 0, one, None, 1, 2048, 3, 11, 4, 0, 1, 0,
                                  *([one] * 7), None)
    assert rc == -2  # hidden too large
    assert lib.consolver_set_step_launch(100, 1) == -2
    assert lib.consolver_set_step_launch(0, 0) == 0
    for code in (0, -1, -2, -3, -4, 700):
        assert lib.consolver_error_string(code)
    with pytest.raises(_lib.ConsolverError):
        _lib.check(-2, "x")


def _cuobjdump():
    import shutil

    return shutil.which("cuobjdump") or ("/usr/local/cuda/bin/cuobjdump"
                                         if os.path.exists("/usr/local/cuda/bin/cuobjdump") else None)


@pytest.mark.skipif(_cuobjdump() is None, reason="cuobjdump not installed")
@pytest.mark.parametrize("obj", ["step_sd.o", "step_fm.o"])
def test_no_load_of_pdl_produced_data_is_hoisted_above_the_wait(obj):
    """The step kernel starts while the policy kernel that writes its coefficient record is still running (programmatic
    dependent launch) and reads the record after griddepcontrol.wait (SASS: ACQBULK).  `__ldg` loads are invariant to
    nvcc, which hoisted two of them above the wait in 12 instantiations (stale coefficients, found by the live
    differential fuzzing).  The loads are now volatile asm (ordinary coherent-path ld.global): in every 128-bit
    instantiation each load before the wait must be one of the 128-bit streaming loads of the model outputs / latent,
    and no load after it may take the non-coherent (LDG.CONSTANT) path."""
    from consolver_b200 import build

    path = os.path.join(build.PKG, "build", obj)
    if not os.path.exists(path):
        build.build_library(force=True)
    sass = subprocess.run([_cuobjdump(), "-sass", path], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    checked = 0
    for fn in funcs:
        name = fn.split("\n", 1)[0].strip()
        if "step_kernel" not in name or "ACQBULK" not in fn:
            continue
        pre, post = fn.split("ACQBULK", 1)
        elems = int(re.search(r"Li(\d+)ELi\d+ELb[01]EEE", name).group(1))      # E: elements per thread-vector
        if elems > 1:
            hoisted = [l for l in re.findall(r"LDG\S*", pre) if ".128" not in l]
            assert not hoisted, f"{name}: {hoisted} before griddepcontrol.wait"
        stale_path = [l for l in re.findall(r"LDG\S*", post) if "CONSTANT" in l]
        assert not stale_path, f"{name}: non-coherent loads after griddepcontrol.wait: {stale_path}"
        checked += 1
    assert checked >= 40
    for src in ("step_kernel.cuh",):
        with open(os.path.join(build.CSRC, src)) as f:
            assert "__ldg(" not in f.read(), f"{src}: __ldg in a kernel that waits on a PDL primary"
