"""The C-ABI shared library: loads (no GPU needed), exports every symbol the header declares, struct layouts of the
ctypes mirror match the C compiler's, and the host-side wrapper fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from drloco_b200 import cabi, lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "drloco_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drl_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(lib.LIB_PATH):
        sys.path.insert(0, REPO)
        import __graft_entry__
        __graft_entry__.build()
    return C.CDLL(lib.LIB_PATH)


def test_library_exports_every_declared_symbol(built_lib):
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/drloco_b200.h but not exported"
    # and the python prototypes cover exactly the declared surface
    assert sorted(lib.PROTOTYPES) == names
    built_lib.drl_version.restype = C.c_int
    assert built_lib.drl_version() == cabi.DRL_ABI_VERSION


def test_struct_layouts_match_the_c_compiler(tmp_path):
    probe = tmp_path / "probe.c"
    probe.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "drloco_b200.h"\n'
                     'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(DrlWalkerModel), sizeof(DrlConfig),'
                     ' offsetof(DrlWalkerModel, dof_body), offsetof(DrlWalkerModel, site_pos),'
                     ' offsetof(DrlConfig, seed), offsetof(DrlConfig, lanes_per_env));return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), str(probe), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(cabi.DrlWalkerModel), C.sizeof(cabi.DrlConfig), cabi.DrlWalkerModel.dof_body.offset,
            cabi.DrlWalkerModel.site_pos.offset, cabi.DrlConfig.seed.offset, cabi.DrlConfig.lanes_per_env.offset]
    assert got == want


def test_enum_tables_match_header():
    src = open(HEADER).read()
    assert int(re.search(r"DRL_EXTRA_COUNT = (\d+)", src).group(1)) == cabi.EXTRA_COUNT == len(cabi.EXTRA_NAMES)
    assert int(re.search(r"DRL_STATS_COUNT = (\d+)", src).group(1)) == cabi.STATS_COUNT == len(cabi.STAT_NAMES)
    for k, name in enumerate(cabi.STAT_NAMES):
        assert re.search(rf"DRL_STAT_{name.upper()} = {k}\b", src), name


def test_argument_errors_do_not_need_a_gpu(built_lib):
    l = lib.load()
    h = C.c_void_p()
    assert l.drl_create(None, C.byref(h)) == -1
    assert b"null" in l.drl_last_error()
    cfg = cabi.DrlConfig()
    cfg.num_envs = 0
    assert l.drl_create(C.byref(cfg), C.byref(h)) == -1
    assert l.drl_step(None, None, None, None, None, None, None, None, None) == -1
    assert l.drl_upload_model(None, None) == -1
    assert l.drl_vecnorm_moments(None, 0, 0, None, None, 0.99, None, None) == -1


def test_no_cpu_fallback():
    """without a CUDA device the product refuses to construct (it must never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from drloco_b200.vec_env import B200MimicVecEnv
    with pytest.raises(lib.DrlError):
        B200MimicVecEnv("StraightMimicWalker", num_envs=4)
    # the package itself never imports the oracle
    pkg = os.path.join(REPO, "drloco_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
