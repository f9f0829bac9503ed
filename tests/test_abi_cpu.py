"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/dlsg.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    path = ge.build()
    return ctypes.CDLL(path)


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'dlsg.h')).read()
    names = sorted(set(re.findall(r'\b(dlsg_[a-z0-9_]+)\s*\(', hdr)))
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header(lib):
    from dlsg import _lib
    hdr = open(os.path.join(ROOT, 'include', 'dlsg.h')).read()
    names = set(re.findall(r'\b(dlsg_[a-z0-9_]+)\s*\(', hdr))
    assert set(_lib.SIGNATURES) == names
    lib.dlsg_sm_arch.restype = ctypes.c_int
    assert lib.dlsg_sm_arch() == 100


def test_struct_sizes_match_c_layout():
    """ctypes mirrors of the parameter structs must have the C compiler's layout (checked by compiling a probe)."""
    import subprocess
    import tempfile
    from dlsg import _lib
    src = '#include <stdio.h>\n#include "dlsg.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(dlsg_gemm_t),sizeof(dlsg_norm_fwd_t),sizeof(dlsg_norm_bwd_t),sizeof(dlsg_lstm_cell_fwd_t),' \
          'sizeof(dlsg_lstm_cell_bwd_t),sizeof(dlsg_softmax_t),sizeof(dlsg_node_attn_fwd_t),sizeof(dlsg_node_attn_bwd_t),sizeof(dlsg_lstm_cell_bwd2_t),sizeof(dlsg_seg_t),sizeof(dlsg_norm_bwd2_t),sizeof(dlsg_adam_seg_t),sizeof(dlsg_region_agg_fwd_t),sizeof(dlsg_region_agg_bwd_t));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 'p.c')
        open(c, 'w').write(src)
        exe = os.path.join(d, 'p')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    ours = [ctypes.sizeof(s) for s in (_lib.GemmT, _lib.NormFwdT, _lib.NormBwdT, _lib.CellFwdT, _lib.CellBwdT, _lib.SoftmaxT,
                                       _lib.AttnFwdT, _lib.AttnBwdT, _lib.CellBwd2T, _lib.SegT, _lib.NormBwd2T, _lib.AdamSegT,
                                       _lib.RegionAggFwdT, _lib.RegionAggBwdT)]
    assert ours == sizes, (ours, sizes)


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from dlsg import ops, _lib
    be = ops.CudaBackend()
    with pytest.raises(_lib.DlsgError):
        be.gemm(torch.zeros(4, 8), torch.zeros(4, 8), torch.zeros(4, 4))
