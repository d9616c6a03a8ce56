"""The C-ABI library loads and exports every symbol include/aewn.h declares; argument validation works without a GPU."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "aewn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(aewn_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from aewn import _lib
    lib = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"libaewn.so does not export {s}"
    assert lib.aewn_version() >= 100


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors must have the C layout: compile-time sizes are baked into a tiny C probe built from the header."""
    import subprocess
    import tempfile
    from aewn import _lib
    src = '#include "aewn.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", ' \
          'sizeof(aewn_act), sizeof(aewn_seg), sizeof(aewn_ntile), sizeof(aewn_tgemm_desc), sizeof(aewn_wgrad_item), ' \
          'sizeof(aewn_wgrad_desc), sizeof(aewn_copy_block), sizeof(aewn_gen_block), sizeof(aewn_gen_desc), ' \
          'sizeof(aewn_wgw_chunk), sizeof(aewn_wgw_unit), sizeof(aewn_wgradw_desc), sizeof(aewn_grcc_fwd_desc), sizeof(aewn_grcc_dgrad_desc), sizeof(aewn_mfcc_desc), sizeof(aewn_act16), sizeof(aewn_wgradh_desc), sizeof(aewn_grcc_gz_desc)); ' \
          'return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "p.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "p"), os.path.join(td, "p.c")],
                       check=True)
        sizes = [int(v) for v in subprocess.run([os.path.join(td, "p")], capture_output=True, text=True).stdout.split()]
    mine = [ctypes.sizeof(t) for t in (_lib.Act, _lib.Seg, _lib.NTile, _lib.TGemmDesc, _lib.WGradItem, _lib.WGradDesc,
                                       _lib.CopyBlock, _lib.GenBlock, _lib.GenDesc, _lib.WGWChunk, _lib.WGWUnit,
                                       _lib.WGradWDesc, _lib.GrccFwdDesc, _lib.GrccDgradDesc, _lib.MfccDesc, _lib.Act16, _lib.WGradHDesc, _lib.GrccGzDesc)]
    assert mine == sizes


def test_invalid_arguments_are_rejected_with_message():
    from aewn import _lib
    lib = _lib.lib()
    rc = lib.aewn_tgemm(None, None)
    assert rc == -1001
    assert b"null descriptor" in lib.aewn_last_error_string()
    d = _lib.TGemmDesc()
    d.n_acts, d.n_segs, d.n_ntiles, d.batch, d.t_begin, d.t_end = 1, 1, 1, 1, 0, 128
    rc = lib.aewn_tgemm(ctypes.byref(d), None)
    assert rc != 0      # null activation pointer or missing driver entry point -- never a crash
    g = _lib.GenDesc()
    g.cluster, g.n_rep, g.n_layers, g.n_blocks = 3, 1, 1, 4
    assert lib.aewn_gen_run(ctypes.byref(g), None) == -1001
    assert b"cluster must be" in lib.aewn_last_error_string()
    g.cluster, g.n_rep = 4, 3
    assert lib.aewn_gen_run(ctypes.byref(g), None) == -1001
    assert b"n_rep must be" in lib.aewn_last_error_string()
