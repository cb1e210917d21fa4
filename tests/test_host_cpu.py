"""CPU-only tests of the boundary and the host logic: the C-ABI library loads and exports every
symbol include/ffcorr.h declares (no compute calls without a GPU), argument errors are reported
through return codes, and the multi-process sharding logic works under gloo (world_size 2)."""
import ctypes
import os
import re
import socket

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ffcorr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ffcorr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from focusflow_official_b200 import _lib

    syms = declared_symbols()
    assert len(syms) >= 12
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/ffcorr.h but not exported"
    assert set(syms) == set(_lib.SYMBOLS), "ctypes prototypes and header disagree"
    assert _lib.lib().ffcorr_version() == 201


def test_ctypes_prototypes_match_the_header_parameter_by_parameter():
    """Every prototype in _lib.SYMBOLS has the arity and the parameter KINDS (pointer / 32-bit int / 64-bit int /
    size_t / float) of its declaration in include/ffcorr.h: a mismatch would silently corrupt arguments."""
    from focusflow_official_b200 import _lib

    text = open(os.path.join(ROOT, "include", "ffcorr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b(?:int|size_t|int64_t|const char\*)\s+(ffcorr_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text)
    assert len(protos) == len(_lib.SYMBOLS)

    def kind_of_c(param):
        p = param.strip()
        if p in ("void", ""):
            return None
        if "*" in p:
            return "ptr"
        if p.startswith("size_t"):
            return "size_t"
        if p.startswith("int64_t"):
            return "i64"
        if p.startswith("float"):
            return "float"
        assert p.startswith("int "), p
        return "int"

    def kind_of_ctypes(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
            return "ptr"
        return {ctypes.c_size_t: "size_t", ctypes.c_int64: "i64", ctypes.c_float: "float", ctypes.c_int: "int"}[t]

    for name, params in protos:
        want = [k for k in (kind_of_c(p) for p in params.split(",")) if k is not None]
        restype, argtypes = _lib._PROTOS[name]
        got = [kind_of_ctypes(t) for t in argtypes]
        assert got == want, (name, got, want)


def test_argument_errors_are_return_codes_not_crashes():
    from focusflow_official_b200 import _lib

    L = _lib.lib()
    # bad shapes are rejected before any CUDA call, so this is safe without a GPU
    assert L.ffcorr_lookup_f32(None, 4, None, None, -1, 8, 8, 4, 1, 0, None) == -1
    assert b"B=-1" in L.ffcorr_last_error()
    ptrs = (ctypes.c_void_p * 4)()
    assert L.ffcorr_lookup_f32(ptrs, 4, 1, 1, 1, 8, 8, 9, 1, 0, None) == -1    # radius out of range
    assert L.ffcorr_lookup_f32(ptrs, 4, 1, 1, 1, 16, 16, 4, 7, 0, None) == -1  # sampler is neither ATEN_CPU nor ATEN_CUDA
    assert b"sampler" in L.ffcorr_last_error()
    assert L.ffcorr_lookup_f32(ptrs, 1, 1, 1, 1, 8192, 4096, 4, 1, 0, None) == -1   # h*w >= 2^24: 32-bit tile offsets
    assert b"2^24" in L.ffcorr_last_error()
    assert L.ffcorr_lookup_f32(ptrs, 4, 1, 1, 1, 4, 4, 4, 1, 0, None) == -1    # map too small for 4 levels
    assert L.ffcorr_pyramid_f32(ptrs, 99, 1, 8, 8, None) == -1
    assert L.ffcorr_volume_f32(1, 1, 1, 1, 0, 8, 8, 0, None, 0, None) == -1    # D = 0
    assert L.ffcorr_volume_workspace_bytes(8, 256, 47, 156, 0) == 2 * 8 * (48 * 160) * 256 * 2 + 256  # padded to 16x16 super-groups; + the fp16 block-scaling words
    assert L.ffcorr_volume_workspace_bytes(8, 256, 47, 156, 3) == 2 * 8 * (48 * 160) * 256 * 4        # tf32: fp32 operands, no scaling
    # tiled geometry: 4x4 tiles, even number of tiles per row
    assert L.ffcorr_tiled_map_elems(47, 156, 0) == 12 * 40 * 16
    assert L.ffcorr_tiled_map_elems(47, 156, 1) == 6 * 20 * 16
    assert L.ffcorr_tiled_map_elems(47, 156, 3) == 2 * 6 * 16
    assert L.ffcorr_volume_workspace_bytes(8, 256, 47, 156, 1) == 0            # fp32 path needs none
    assert L.ffcorr_pwc81_f32(1, 1, 1, 1, 0, 4, 4, -1.0, None) == -1
    # the tiled / fused / chunked / scaled entry points validate before touching the device too
    assert L.ffcorr_tiled_supported(4, 47, 156) == 1 and L.ffcorr_tiled_supported(5, 47, 156) == 0
    assert L.ffcorr_tiled_supported(4, 100, 160) == 1                           # fused build: no shared-memory map limit
    assert L.ffcorr_build_tiled_f32(1, 1, None, 4, 1, 256, 16, 16, 0, None, 0, None) == -1          # null level table
    assert L.ffcorr_build_tiled_f32(1, 1, ptrs, 5, 1, 256, 64, 64, 0, None, 0, None) == -1          # > 4 levels
    assert L.ffcorr_build_tiled_chunk_f32(ptrs, 4, 1, 256, 16, 16, 0, 32, 0, None, 0, None) == -1   # null level pointers
    assert L.ffcorr_lookup_tiled_chunk_f32(ptrs, 4, 1, 1, 1, 16, 16, 250, 32, 4, 1, 0, None) == -1  # chunk beyond the map
    assert b"query range" in L.ffcorr_last_error()
    assert L.ffcorr_volume_scaled_f32(1, 1, 1, 1, 8, 8, 8, 0, ctypes.c_float(0.0), None, 0, None) == -1
    assert b"divisor" in L.ffcorr_last_error()
    assert L.ffcorr_stage_operands_f32(1, 1, 9, 1, 8, 8, 8, 0, None, 0, None) == -1                 # 9 levels
    # round-2 entry points: fp16 storage, grouped tile order, backwarp
    assert L.ffcorr_build_tiled_f16(1, 1, ptrs, 1, 1, 256, 16, 16, 0, None, 0, None) == -1           # needs 2-4 levels
    assert L.ffcorr_build_tiled_f16(1, 1, ptrs, 4, 1, 256, 16, 16, 1, None, 0, None) == -1           # no fp32-operand variant
    assert L.ffcorr_build_grouped_f32(1, 1, ptrs, 5, 1, 256, 64, 64, 0, None, 0, None) == -1         # > 4 levels
    assert L.ffcorr_lookup_tiled_f16(ptrs, 4, 1, 1, 1, 16, 16, 4, 1, 0, None) == -1                  # NCHW output: channels-last only
    assert L.ffcorr_grouped_level_elems(47, 156, 0, 8, 7332) == 8 * 230 * 12 * 40 * 512
    assert L.ffcorr_grouped_level_elems(47, 156, 3, 1, 33) == 2 * 2 * 6 * 512
    assert L.ffcorr_backwarp_f32(1, 1, 1, 1, 1, 1, 0, 8, 8, ctypes.c_float(1.0), None) == -1         # C = 0
    assert L.ffcorr_backwarp_f32(None, None, None, None, None, 0, 32, 8, 8, ctypes.c_float(1.0), None) == 0   # empty batch
    # the lookup fused with convc1 (update.py:82-83,90): built for 4 levels x radius 4 and Conv2d(324, 256, 1)
    assert L.ffcorr_convc1_packed_bytes() == 256 * 384 * 2
    assert L.ffcorr_lookup_convc1_tiled_f32(ptrs, 3, 1, 1, 1, 1, 1, 1, 16, 16, 4, 1, None) == -1
    assert b"4 levels x radius 4" in L.ffcorr_last_error()
    assert L.ffcorr_lookup_convc1_tiled_f32(ptrs, 4, 1, 1, 1, 1, 1, 1, 16, 16, 3, 1, None) == -1
    assert L.ffcorr_lookup_convc1_tiled_f32(ptrs, 4, 1, None, 1, 1, 1, 1, 16, 16, 4, 1, None) == -1     # null packed weight
    assert L.ffcorr_lookup_convc1_tiled_f32(ptrs, 4, 1, 1, 1, 1, None, 1, 16, 16, 4, 1, None) == -1     # null tile counter
    assert L.ffcorr_lookup_convc1_tiled_f32(ptrs, 4, 16, 16, 16, 16, 16, 1, 16, 16, 4, 1, None) == -1   # null level pointers
    assert L.ffcorr_lookup_convc1_tiled_f32(None, 4, None, None, None, None, None, 0, 16, 16, 4, 1, None) == 0
    assert L.ffcorr_pack_convc1_weight(1, 128, 324, 1, None) == -1
    assert b"Conv2d(324, 256, 1)" in L.ffcorr_last_error()
    # empty batches are a no-op, even with null pointers
    assert L.ffcorr_lookup_f32(None, 4, None, None, 0, 16, 16, 4, 1, 1, None) == 0
    assert L.ffcorr_build_tiled_f32(None, None, ptrs, 4, 0, 256, 16, 16, 0, None, 0, None) == 0
    assert L.ffcorr_lookup_tiled_chunk_f32(ptrs, 4, None, None, 0, 16, 16, 0, 32, 4, 0, 1, None) == 0
    with pytest.raises(RuntimeError):
        _lib.check(-1, "x")


def test_sampler_default_and_validation():
    """The library has no global switch any more: the sampler travels with every call; Python keeps only a default."""
    import focusflow_official_b200 as ff
    from focusflow_official_b200 import _lib

    assert ff.get_sampler_semantics() == "cuda"          # what a GPU user of the reference gets
    ff.set_sampler_semantics("cpu")
    assert ff.get_sampler_semantics() == "cpu"
    ff.set_sampler_semantics("cuda")
    with pytest.raises(ValueError):
        ff.set_sampler_semantics("gpu")
    assert not hasattr(_lib.lib(), "ffcorr_set_sampler_semantics")


def test_cpu_tensors_are_refused_not_silently_computed():
    import focusflow_official_b200 as ff

    f = torch.zeros(1, 8, 16, 16)
    with pytest.raises(NotImplementedError):
        ff.CorrBlock(f, f)
    with pytest.raises(NotImplementedError):
        ff.FunctionCorrelation(f, f)
    g = ff.coords_grid(2, 3, 5, "cpu")
    assert g.shape == (2, 2, 3, 5) and float(g[1, 0, 2, 4]) == 4.0 and float(g[1, 1, 2, 4]) == 2.0


def test_missing_extension_fails_loudly_in_a_fresh_process():
    """No fallback: with the shared library absent, the first use of the package raises and names the build command."""
    import subprocess
    import sys

    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from focusflow_official_b200 import _lib\n"
        "_lib.LIB_PATH = _lib.LIB_PATH + '.absent'\n"
        "import focusflow_official_b200 as ff\n"
        "try:\n"
        "    ff.CorrBlock  # importing is fine; the first USE must fail\n"
        "    _lib.lib()\n"
        "except RuntimeError as e:\n"
        "    assert 'missing' in str(e) and 'make -C' in str(e), str(e)\n"
        "    print('LOUD')\n" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "LOUD" in p.stdout, p.stdout + p.stderr[-1500:]


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "focusflow_official_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "oracle/" not in src.replace("oracle/__init__", ""), fn


def test_shard_range():
    from focusflow_official_b200.sharding import shard_range

    for total in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(64, 3, 8) == (24, 32)  # config 4: 64 pairs over 8 GPUs
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from focusflow_official_b200.sharding import job_throughput, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(9, rank, world)
    seconds = 1.0 + rank  # rank 1 is the slow one
    thr = job_throughput(b - a, seconds, world)
    sizes = [None] * world
    dist.all_gather_object(sizes, (a, b))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, thr, sizes))


def test_sharding_under_gloo_world2():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, thr, sizes in res:
        assert sizes == [(0, 5), (5, 9)]
        assert abs(thr - 9 / 2.0) < 1e-9  # sum of pairs / max time over ranks


def test_bench_reference_arm_prints_one_contract_line():
    """`python bench.py --impl reference` (the CPU arm the driver runs next to ours): exactly ONE line on stdout, the
    contract keys, the reference-arm extras, and no GPU needed."""
    import json
    import subprocess
    import sys

    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    from oracle import reference_loader as RL

    # the unmodified reference when its sources are installed (/root/reference or baseline/_ref), else the port
    assert d["cpu_baseline"]["kind"] == ("reference" if RL.reference_root() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"] and d["steps"] >= 5
