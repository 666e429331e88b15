"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/evdb.h declares, and refuses to work (loudly) without a B200 -- no compute calls."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "evdb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(evdb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from erlvectordb_b200 import _native as N
    L = N.lib()
    syms = declared_symbols()
    assert len(syms) >= 24
    bound = {name for name, _, _ in N.SYMBOLS}
    for s in syms:
        assert hasattr(L, s), f"libevdb_b200.so does not export {s}"
        assert s in bound, f"_native.py does not bind {s}"
    assert L.evdb_abi_version() == 2   # EVDB_ABI_VERSION: n_shards / devices in evdb_opts, the ABI-2 tail of evdb_stats


def test_error_strings_map_to_reference_atoms():
    from erlvectordb_b200 import _native as N
    L = N.lib()
    assert L.evdb_strerror(N.E_DIM_MISMATCH) == b"dimension_mismatch"   # vector_store.erl:221
    assert L.evdb_strerror(N.E_BAD_VECTOR) == b"invalid_vector_format"  # vector_store.erl:216,222,225
    assert L.evdb_strerror(0) == b"ok"


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from erlvectordb_b200 import _native as N
    from erlvectordb_b200 import vector_store, vector_compression
    assert N.lib().evdb_init(None, 0) == N.E_NO_DEVICE
    with pytest.raises(N.EvdbError):
        vector_store.start_link("no_gpu_store")
    with pytest.raises(N.EvdbError):
        vector_compression.compress_vector([1.0, 2.0], "quantization_8bit")


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or call it."""
    pkg = os.path.join(ROOT, "erlvectordb_b200")
    banned = ("import oracle", "from oracle", "libevdb_oracle", "evo_", "oracle.")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                src = open(os.path.join(dirpath, f)).read()
                for b in banned:
                    assert b not in src, f"{f} references the oracle ({b!r})"


def test_validate_vector_mirror():
    from erlvectordb_b200.vector_store import validate_vector, term_key
    assert validate_vector([1.0, 2, 3.5], None) == ("ok", 3)          # ints are numbers
    assert validate_vector([1.0, 2.0], 3) == ("error", "dimension_mismatch")
    assert validate_vector([1.0, "a"], 2) == ("error", "invalid_vector_format")
    assert validate_vector("abc", None) == ("error", "invalid_vector_format")
    assert validate_vector([True, 1.0], 2) == ("error", "invalid_vector_format")  # atoms are not numbers
    assert validate_vector(np.zeros(4, dtype=np.float32), 4) == ("ok", 4)
    # Erlang term order: binaries bytewise then by length; numbers < atoms < binaries
    assert term_key(b"v1") < term_key(b"v2") < term_key(b"v2a")
    assert term_key(3) < term_key("atom") < term_key(b"bin")


def test_erlang_nif_shim_compiles_against_the_abi(tmp_path):
    """erlang/c_src/evdb_nif.c (the NIF a maintainer adds) must at least parse and
    type-check against include/evdb.h.  No Erlang/OTP exists in this image, so erl_nif.h is the
    declarations-only stub kept next to the shim."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    shutil.copy(os.path.join(root, "erlang", "c_src", "erl_nif_stub.h"), tmp_path / "erl_nif.h")
    r = subprocess.run([gcc, "-fsyntax-only", "-Wall", "-Werror=implicit-function-declaration",
                        f"-I{tmp_path}", f"-I{os.path.join(root, 'include')}",
                        os.path.join(root, "erlang", "c_src", "evdb_nif.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of evdb_opts / evdb_stats / evdb_search_opts (and therefore the NIF, which uses the
    C structs directly) must agree with include/evdb.h field by field: a C program prints the sizes and
    offsets the compiler assigns, ctypes must assign the same."""
    import ctypes as C
    import shutil
    import subprocess
    from erlvectordb_b200 import _native as N
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    structs = {"evdb_opts": N.Opts, "evdb_stats": N.Stats, "evdb_search_opts": N.SearchOpts}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "evdb.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  printf("EVDB_MAX_SHARDS %d\\n", EVDB_MAX_SHARDS);', '  printf("EVDB_ABI_VERSION %d\\n", EVDB_ABI_VERSION);',
              '  return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    r = subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    assert int(got["EVDB_MAX_SHARDS"]) == 16 == len(N.Opts().devices)
    assert int(got["EVDB_ABI_VERSION"]) == N.lib().evdb_abi_version()


def test_multi_device_slot_arithmetic():
    """Host arithmetic of the one-handle multi-device store (csrc/mstore.cu): global slot g lives on shard
    g mod S at local slot g div S; appends and swap-with-last deletes keep every shard dense and within one
    row of the others.  Replayed here on plain lists against a flat model."""
    import random
    rnd = random.Random(5)
    for S in (2, 3, 8):
        shards = [[] for _ in range(S)]
        flat = []
        for step in range(2000):
            if not flat or rnd.random() < 0.6:
                g = len(flat)
                assert len(shards[g % S]) == g // S          # an append lands at the end of its shard
                shards[g % S].append(step)
                flat.append(step)
            else:
                g, last = rnd.randrange(len(flat)), len(flat) - 1
                assert len(shards[last % S]) - 1 == last // S   # the global last row is the last row of its shard
                shards[g % S][g // S] = shards[last % S][last // S]
                shards[last % S].pop()
                flat[g] = flat[last]
                flat.pop()
            n = len(flat)
            assert [len(sh) for sh in shards] == [(n - j + S - 1) // S if n > j else 0 for j in range(S)]
            assert max(map(len, shards)) - min(map(len, shards)) <= 1
        assert all(shards[g % S][g // S] == flat[g] for g in range(len(flat)))
