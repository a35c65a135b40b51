"""CPU: libcppf_b200.so loads (cross-compiled for sm_100a) and exports every symbol
include/cppf_b200.h declares; the ctypes table mirrors the header."""
import os
import re

from conftest import ROOT
from cppf_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cppf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cppf_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in cppf_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in cppf_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_host_only_entry_points():
    L = _lib.lib()
    assert L.cppf_abi_version() >= 1
    assert L.cppf_ppf_feat_dim() == 40
    assert L.cppf_ppf_blob_floats(141) == 12336
    assert L.cppf_compact_scratch_bytes(100000) > 0
    assert L.cppf_launch_count() == 0 or L.cppf_launch_count() > 0


def test_pose_args_struct_mirror_matches_the_compiled_layout():
    import ctypes as C
    L = _lib.lib()
    assert C.sizeof(_lib.PoseArgs) == L.cppf_pose_args_bytes()
    assert L.cppf_pose_record_doubles() == 16
    assert L.cppf_timing_stages() == 10 and L.cppf_timing_stage_name(3) == b"encode_sample"
    assert L.cppf_pose_workspace_bytes(4096, 0, 60, 17820, 0, 72, 480, 10000) > 5 * 4096 * 4096
    assert L.cppf_vote_routed_supported(64, 64, 64) == 1 and L.cppf_vote_routed_supported(640, 640, 640) == 0


def test_missing_library_fails_loudly_and_nothing_falls_back():
    """No CPU / PyTorch fallback: without the built .so the binding raises on first use."""
    import subprocess
    import sys
    code = ("import os; os.environ['CPPF_B200_LIB'] = '/nonexistent/libcppf_b200.so'\n"
            "from cppf_b200 import _lib\n"
            "try:\n"
            "    _lib.lib()\n"
            "except RuntimeError as e:\n"
            "    assert 'missing' in str(e) and 'no CPU or PyTorch fallback' in str(e); print('raised')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "raised" in out.stdout, out.stderr


def test_dropin_package_shadows_the_reference_module_paths():
    """`from models.model import PPFEncoder, PointEncoder` / `from models.voting import ...` (nocs/inference.py:3,17)
    resolve to this implementation when <repo>/dropin is ahead on sys.path; constructors take the reference's kwargs."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from models.model import PPFEncoder, PointEncoder\n"
            "from models.voting import rot_voting_kernel, backvote_kernel, ppf_kernel, findpeak_kernel\n"
            "import cppf_b200.model as m\n"
            "assert PPFEncoder is m.PPFEncoder and PointEncoder is m.PointEncoder\n"
            "pe = PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32)\n"
            "ppf = PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141)\n"
            "assert 'final.weight' in ppf.state_dict() and any(k.startswith('spconvs.0.') for k in pe.state_dict())\n"
            "assert all(callable(k) for k in (rot_voting_kernel, backvote_kernel, ppf_kernel, findpeak_kernel))\n"
            "print('ok')\n") % (root, os.path.join(root, "dropin"))
    out = subprocess.run([sys.executable, "-c", code], cwd="/tmp", capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_numpy_mirror_of_pose_args_matches_the_compiled_layout():
    """enqueue_batch fills a whole batch of cppf_pose_args column-wise through a numpy structured dtype."""
    from cppf_b200 import pipeline
    dt = pipeline._args_dtype()
    assert dt.itemsize == _lib.lib().cppf_pose_args_bytes()
    assert dt.names == tuple(f for f, _ in _lib.PoseArgs._fields_)
    assert dt.fields["sample_pairs"][1] == _lib.PoseArgs.sample_pairs.offset
