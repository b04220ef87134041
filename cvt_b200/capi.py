"""ctypes binding of the C ABI (include/b200nn.h) -- the same entry points a cgo/JNI/C++ caller
binds.  Used by the tests and bench.py; numpy arrays are HOST buffers, `*_dev` methods take raw
device pointers (e.g. torch tensors' data_ptr()).

There is no fallback: if the CUDA library has not been built, or no B200 is present, loading /
context creation raises."""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libb200nn.so")

# every symbol include/b200nn.h declares (tests check that the built library exports them all)
SYMBOLS = [
    "b200nn_last_error", "b200nn_version", "b200nn_ctx_create", "b200nn_ctx_destroy", "b200nn_ctx_set_stream",
    "b200nn_ctx_synchronize", "b200nn_ctx_launch_count", "b200nn_ctx_event_record", "b200nn_ctx_event_elapsed_ms",
    "b200nn_flat_create", "b200nn_flat_destroy", "b200nn_flat_add", "b200nn_flat_remove", "b200nn_flat_size",
    "b200nn_flat_search", "b200nn_flat_search_dev", "b200nn_flat_save", "b200nn_flat_load", "b200nn_flat_load_hnsw", "b200nn_flat_info",
    "b200nn_pq_create", "b200nn_pq_load_model", "b200nn_pq_destroy", "b200nn_pq_set_clamp", "b200nn_pq_info",
    "b200nn_pq_rotate", "b200nn_pq_encode", "b200nn_pq_add", "b200nn_pq_add_dev", "b200nn_pq_add_rotated", "b200nn_pq_get_rows",
    "b200nn_pq_build_lut", "b200nn_pq_scores", "b200nn_pq_query_groups", "b200nn_pq_search", "b200nn_pq_search_dev", "b200nn_topk_merge_dev",
    "b200nn_topk_merge_grid_dev", "b200nn_pq_scan_plan",
    "b200nn_pq_save_index", "b200nn_pq_save_index_n", "b200nn_pq_load_index", "b200nn_pq_last_timing", "b200nn_pq_scan_bytes",
    "b200nn_pq_scores_dev", "b200nn_pq_append_coded",
    "b200nn_comm_get_unique_id", "b200nn_comm_create", "b200nn_comm_destroy", "b200nn_comm_info", "b200nn_pq_search_sharded_dev",
    "b200nn_plan_layout", "b200nn_mpq_create", "b200nn_mpq_load_model", "b200nn_mpq_destroy", "b200nn_mpq_info", "b200nn_mpq_shard_rows", "b200nn_mpq_set_clamp",
    "b200nn_mpq_rotate", "b200nn_mpq_add", "b200nn_mpq_add_rotated", "b200nn_mpq_search", "b200nn_mpq_scores", "b200nn_mpq_save_index",
    "b200nn_mpq_load_index",
    "b200nn_sq_create", "b200nn_sq_destroy", "b200nn_sq_train_minmax", "b200nn_sq_encode", "b200nn_sq_decode",
    "b200nn_sq_encode_dev",
    "b200nn_proj_create", "b200nn_proj_load_model", "b200nn_pca_read_model", "b200nn_proj_destroy", "b200nn_proj_apply", "b200nn_proj_apply_dev", "b200nn_rootsift", "b200nn_rootsift_dev",
    "b200nn_kmeans", "b200nn_kmeans_init_rows", "b200nn_kmeans_plan_update", "b200nn_pq_train", "b200nn_pq_write_model",
]

_lib = None


class B200nnError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library; raises if it was not built (python -m cvt_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200nnError(f"{LIB_PATH} is missing: build it with `python -m cvt_b200.build` (no CPU fallback exists)")
        _lib = C.CDLL(LIB_PATH)
        _lib.b200nn_last_error.restype = C.c_char_p
        _lib.b200nn_version.restype = C.c_char_p
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise B200nnError(f"{what} failed ({rc}): {load().b200nn_last_error().decode(errors='replace')}")


def _vp(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Context:
    def __init__(self, device: int = 0):
        self.h = C.c_void_p()
        _check(load().b200nn_ctx_create(C.c_int(device), C.byref(self.h)), "ctx_create")
        self.device = device
        self._children = []  # weakrefs: indexes must be destroyed before their context

    def _adopt(self, child):
        self._children.append(weakref.ref(child))

    def close(self):
        if self.h:
            for r in self._children:
                ch = r()
                if ch is not None:
                    ch.close()
            self._children = []
            load().b200nn_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int | None):
        _check(load().b200nn_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)), "ctx_set_stream")

    def synchronize(self):
        _check(load().b200nn_ctx_synchronize(self.h), "ctx_synchronize")

    def launch_count(self) -> int:
        v = C.c_uint64()
        _check(load().b200nn_ctx_launch_count(self.h, C.byref(v)), "ctx_launch_count")
        return int(v.value)

    def event_record(self, slot: int):
        _check(load().b200nn_ctx_event_record(self.h, C.c_int(slot)), "ctx_event_record")

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        _check(load().b200nn_ctx_event_elapsed_ms(self.h, C.c_int(a), C.c_int(b), C.byref(ms)), "ctx_event_elapsed_ms")
        return float(ms.value)

    def topk_merge_dev(self, keys_dev: int, L: int, nq: int, k: int, out_dist_dev: int, out_id_dev: int):
        _check(load().b200nn_topk_merge_dev(self.h, C.c_void_p(keys_dev), C.c_int(L), C.c_size_t(nq), C.c_size_t(k),
                                            C.c_void_p(out_dist_dev), C.c_void_p(out_id_dev)), "topk_merge_dev")


    def topk_merge_grid_dev(self, keys_dev: int, n_chunks: int, L: int, chunk_q: int, nq: int, k: int, out_dist_dev: int,
                            out_id_dev: int):
        _check(load().b200nn_topk_merge_grid_dev(self.h, C.c_void_p(keys_dev), C.c_int(n_chunks), C.c_int(L), C.c_size_t(chunk_q),
                                                 C.c_size_t(nq), C.c_size_t(k), C.c_void_p(out_dist_dev), C.c_void_p(out_id_dev)),
               "topk_merge_grid_dev")


def scan_plan(sm_count: int, M: int, nq: int, n_rows: int):
    """Host-only: the fused scan's work plan -> (n_full, n_tail, slices, desc[n_tail, 2, 4])."""
    nf, nt, sl = C.c_int(), C.c_int(), C.c_int()
    desc = np.zeros((max(1, sm_count), 2, 4), dtype=np.int32)
    _check(load().b200nn_pq_scan_plan(C.c_int(sm_count), C.c_int(M), C.c_size_t(nq), C.c_size_t(n_rows), C.byref(nf), C.byref(nt),
                                      C.byref(sl), _vp(desc), C.c_size_t(desc.size)), "pq_scan_plan")
    return nf.value, nt.value, sl.value, desc[:nt.value]


class PQIndex:
    """IVFOPQ drop-in surface (opq/src/IVFOPQ.h:31-47) over the C ABI."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle
        ctx._adopt(self)
        D, K, M, ks = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(load().b200nn_pq_info(self.h, C.byref(D), C.byref(K), C.byref(M), C.byref(ks), None, None), "pq_info")
        self.D, self.K, self.M, self.ksub = D.value, K.value, M.value, ks.value

    @classmethod
    def create(cls, ctx, coarse, codebooks, perm=None, R=None, clamp=1.0):
        coarse, codebooks = _f32(coarse), _f32(codebooks)
        K, D = coarse.shape
        M, ksub, ds = codebooks.shape
        assert M * ds == D
        perm_a = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
        R_a = None if R is None else _f32(R)
        h = C.c_void_p()
        _check(load().b200nn_pq_create(ctx.h, C.c_int(D), C.c_int(K), C.c_int(M), C.c_int(ksub), _vp(coarse), _vp(codebooks),
                                       _vp(perm_a), _vp(R_a), C.c_float(clamp), C.byref(h)), "pq_create")
        return cls(ctx, h)

    @classmethod
    def load_model(cls, ctx, path: str):
        h = C.c_void_p()
        _check(load().b200nn_pq_load_model(ctx.h, path.encode(), C.byref(h)), "pq_load_model")
        return cls(ctx, h)

    @classmethod
    def load_index(cls, ctx, path: str, perm=None, clamp=1.0):
        perm_a = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
        h = C.c_void_p()
        _check(load().b200nn_pq_load_index(ctx.h, path.encode(), _vp(perm_a), C.c_float(clamp), C.byref(h)), "pq_load_index")
        return cls(ctx, h)

    def close(self):
        if self.h:
            load().b200nn_pq_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_clamp(self, clamp: float):
        _check(load().b200nn_pq_set_clamp(self.h, C.c_float(clamp)), "pq_set_clamp")

    @property
    def n_rows(self) -> int:
        v = C.c_uint64()
        _check(load().b200nn_pq_info(self.h, None, None, None, None, C.byref(v), None), "pq_info")
        return int(v.value)

    @property
    def n_groups(self) -> int:
        v = C.c_uint64()
        _check(load().b200nn_pq_info(self.h, None, None, None, None, None, C.byref(v)), "pq_info")
        return int(v.value)

    def rotate(self, x):
        x = _f32(x)
        y = np.empty_like(x)
        _check(load().b200nn_pq_rotate(self.h, _vp(x), C.c_size_t(x.shape[0]), _vp(y)), "pq_rotate")
        return y

    def encode(self, x_rot):
        x_rot = _f32(x_rot)
        n = x_rot.shape[0]
        lists = np.empty(n, dtype=np.int32)
        codes = np.empty((n, self.M), dtype=np.uint8)
        _check(load().b200nn_pq_encode(self.h, _vp(x_rot), C.c_size_t(n), _vp(lists), _vp(codes)), "pq_encode")
        return lists, codes

    def add(self, x_raw, group_ids=None):
        x_raw = _f32(x_raw)
        g = None if group_ids is None else np.ascontiguousarray(group_ids, dtype=np.int32)
        _check(load().b200nn_pq_add(self.h, _vp(x_raw), C.c_size_t(x_raw.shape[0]), _vp(g)), "pq_add")

    def add_rotated(self, x_rot, group_ids=None):
        x_rot = _f32(x_rot)
        g = None if group_ids is None else np.ascontiguousarray(group_ids, dtype=np.int32)
        _check(load().b200nn_pq_add_rotated(self.h, _vp(x_rot), C.c_size_t(x_rot.shape[0]), _vp(g)), "pq_add_rotated")

    def add_dev(self, x_dev_ptr: int, n: int, group_dev_ptr: int | None = None):
        _check(load().b200nn_pq_add_dev(self.h, C.c_void_p(x_dev_ptr), C.c_size_t(n), C.c_void_p(group_dev_ptr or 0)), "pq_add_dev")

    def get_rows(self, start=0, n=None):
        n = self.n_rows - start if n is None else n
        lists = np.empty(n, dtype=np.int32)
        groups = np.empty(n, dtype=np.int32)
        codes = np.empty((n, self.M), dtype=np.uint8)
        _check(load().b200nn_pq_get_rows(self.h, C.c_uint64(start), C.c_size_t(n), _vp(lists), _vp(groups), _vp(codes)), "pq_get_rows")
        return lists, groups, codes

    def build_lut(self, q_rot, nprobe=1):
        q_rot = _f32(q_rot)
        nq = q_rot.shape[0]
        lists = np.empty((nq, nprobe), dtype=np.int32)
        lut = np.empty((nq, nprobe, self.M, self.ksub), dtype=np.float32)
        _check(load().b200nn_pq_build_lut(self.h, _vp(q_rot), C.c_size_t(nq), C.c_int(nprobe), _vp(lists), _vp(lut)), "pq_build_lut")
        return lists, lut

    def scores(self, q_raw, nprobe=3):
        q_raw = _f32(q_raw)
        out = np.empty((q_raw.shape[0], self.n_groups), dtype=np.float32)
        _check(load().b200nn_pq_scores(self.h, _vp(q_raw), C.c_size_t(q_raw.shape[0]), C.c_int(nprobe), _vp(out)), "pq_scores")
        return out

    def query_groups(self, q_raw, frame_off, k, nprobe=3):
        """Multi-frame query videos -> the k best indexed videos each (multi_frame_index_test.cpp:54-68)."""
        q_raw = _f32(q_raw)
        off = np.ascontiguousarray(frame_off, dtype=np.int64)
        nv = off.shape[0] - 1
        S = np.empty((nv, k), dtype=np.float32)
        G = np.empty((nv, k), dtype=np.uint64)
        _check(load().b200nn_pq_query_groups(self.h, _vp(q_raw), _vp(off), C.c_size_t(nv), C.c_int(nprobe), C.c_size_t(k), _vp(S), _vp(G)),
               "pq_query_groups")
        return S, G

    def search(self, q_raw, k, nprobe=1):
        q_raw = _f32(q_raw)
        nq = q_raw.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.uint64)
        _check(load().b200nn_pq_search(self.h, _vp(q_raw), C.c_size_t(nq), C.c_int(nprobe), C.c_size_t(k), _vp(D), _vp(I)), "pq_search")
        return D, I

    def search_host_ptr(self, q_ptr: int, nq: int, k: int, nprobe: int, out_dist_ptr: int, out_id_ptr: int):
        """Same call as search() with caller-owned (e.g. pinned) HOST buffers."""
        _check(load().b200nn_pq_search(self.h, C.c_void_p(q_ptr), C.c_size_t(nq), C.c_int(nprobe), C.c_size_t(k),
                                       C.c_void_p(out_dist_ptr), C.c_void_p(out_id_ptr)), "pq_search")

    def search_dev(self, q_dev_ptr: int, nq: int, k: int, nprobe: int, out_dist_ptr: int, out_id_ptr: int, out_key_ptr: int = 0,
                   id_base: int = 0):
        _check(load().b200nn_pq_search_dev(self.h, C.c_void_p(q_dev_ptr), C.c_size_t(nq), C.c_int(nprobe), C.c_size_t(k),
                                           C.c_void_p(out_dist_ptr), C.c_void_p(out_id_ptr), C.c_void_p(out_key_ptr),
                                           C.c_uint64(id_base)), "pq_search_dev")

    def search_sharded_dev(self, comm: "Comm", row_shards: int, q_dev_ptr: int, nq: int, k: int, nprobe: int, id_base: int,
                           out_dist_ptr: int, out_id_ptr: int):
        """One step of the sharded search on this rank (collective over comm): local scan of this rank's query chunk ->
        one ncclAllGather of the per-rank records -> merge.  q_dev_ptr = the whole batch."""
        _check(load().b200nn_pq_search_sharded_dev(self.h, comm.h, C.c_int(row_shards), C.c_void_p(q_dev_ptr), C.c_size_t(nq), C.c_int(nprobe),
                                                   C.c_size_t(k), C.c_uint64(id_base), C.c_void_p(out_dist_ptr), C.c_void_p(out_id_ptr)),
               "pq_search_sharded_dev")

    def append_coded(self, lists, codes, group_ids=None):
        lists = np.ascontiguousarray(lists, dtype=np.int32)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        g = None if group_ids is None else np.ascontiguousarray(group_ids, dtype=np.int32)
        _check(load().b200nn_pq_append_coded(self.h, _vp(lists), _vp(g), _vp(codes), C.c_size_t(lists.shape[0])), "pq_append_coded")

    def last_timing(self):
        ms = (C.c_float * 4)()
        _check(load().b200nn_pq_last_timing(self.h, ms), "pq_last_timing")
        return dict(rotate_ms=ms[0], lut_ms=ms[1], scan_ms=ms[2], merge_ms=ms[3])

    def save_index(self, dir_or_path: str, group_paths=None):
        arr = None
        if group_paths is not None:
            arr = (C.c_char_p * len(group_paths))(*[p.encode() for p in group_paths])
        _check(load().b200nn_pq_save_index(self.h, dir_or_path.encode(), arr), "pq_save_index")


COMM_ID_BYTES = 128


def plan_layout(n_ranks: int, n_rows: int, batch: int, M: int, k: int, sm_count: int = 148):
    """(row shards R, query chunks Q) of the rank grid, from the library's cost model (host-only)."""
    r, q = C.c_int(), C.c_int()
    _check(load().b200nn_plan_layout(C.c_int(n_ranks), C.c_uint64(n_rows), C.c_uint64(batch), C.c_int(M), C.c_int(k), C.c_int(sm_count),
                                     C.byref(r), C.byref(q)), "plan_layout")
    return r.value, q.value


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 draws it, every rank passes it to Comm)."""
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    _check(load().b200nn_comm_get_unique_id(buf), "comm_get_unique_id")
    return bytes(buf)


class Comm:
    """One rank of a multi-GPU job: NCCL communicator (ncclCommInitRank) + exchange buffers, owned by the C library."""

    def __init__(self, ctx: Context, rank: int, nranks: int, unique_id: bytes | None):
        self.h = C.c_void_p()
        idb = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(unique_id) if unique_id is not None else None
        _check(load().b200nn_comm_create(ctx.h, C.c_int(rank), C.c_int(nranks), idb, C.byref(self.h)), "comm_create")
        self.ctx, self.rank, self.nranks = ctx, rank, nranks
        ctx._adopt(self)

    def nccl_version(self) -> int:
        v = C.c_int()
        _check(load().b200nn_comm_info(self.h, None, None, C.byref(v)), "comm_info")
        return int(v.value)

    def close(self):
        if self.h:
            load().b200nn_comm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiPQ:
    """IVFOPQ drop-in row-sharded over several GPUs of THIS process (b200nn_mpq_*): host buffers in and out."""

    def __init__(self, handle):
        self.h = handle
        D, K, M, ks, nd, pe = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(load().b200nn_mpq_info(self.h, C.byref(D), C.byref(K), C.byref(M), C.byref(ks), None, None, C.byref(nd), C.byref(pe)), "mpq_info")
        self.D, self.K, self.M, self.ksub, self.n_devices, self.peer_exchange = D.value, K.value, M.value, ks.value, nd.value, bool(pe.value)

    @staticmethod
    def _devs(devices):
        d = np.ascontiguousarray(devices, dtype=np.int32)
        return d, C.c_int(d.shape[0])

    @classmethod
    def create(cls, devices, coarse, codebooks, perm=None, R=None, clamp=1.0):
        coarse, codebooks = _f32(coarse), _f32(codebooks)
        K, D = coarse.shape
        M, ksub, ds = codebooks.shape
        perm_a = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
        R_a = None if R is None else _f32(R)
        d, nd = cls._devs(devices)
        h = C.c_void_p()
        _check(load().b200nn_mpq_create(_vp(d), nd, C.c_int(D), C.c_int(K), C.c_int(M), C.c_int(ksub), _vp(coarse), _vp(codebooks), _vp(perm_a),
                                        _vp(R_a), C.c_float(clamp), C.byref(h)), "mpq_create")
        return cls(h)

    @classmethod
    def load_model(cls, devices, path: str):
        d, nd = cls._devs(devices)
        h = C.c_void_p()
        _check(load().b200nn_mpq_load_model(_vp(d), nd, path.encode(), C.byref(h)), "mpq_load_model")
        return cls(h)

    @classmethod
    def load_index(cls, devices, path: str, perm=None, clamp=1.0):
        d, nd = cls._devs(devices)
        perm_a = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
        h = C.c_void_p()
        _check(load().b200nn_mpq_load_index(_vp(d), nd, path.encode(), _vp(perm_a), C.c_float(clamp), C.byref(h)), "mpq_load_index")
        return cls(h)

    def close(self):
        if self.h:
            load().b200nn_mpq_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _info64(self):
        n, g = C.c_uint64(), C.c_uint64()
        _check(load().b200nn_mpq_info(self.h, None, None, None, None, C.byref(n), C.byref(g), None, None), "mpq_info")
        return int(n.value), int(g.value)

    @property
    def n_rows(self):
        return self._info64()[0]

    @property
    def n_groups(self):
        return self._info64()[1]

    def shard_rows(self):
        r = np.zeros(self.n_devices, dtype=np.uint64)
        _check(load().b200nn_mpq_shard_rows(self.h, _vp(r)), "mpq_shard_rows")
        return r.astype(np.int64)

    def set_clamp(self, clamp: float):
        _check(load().b200nn_mpq_set_clamp(self.h, C.c_float(clamp)), "mpq_set_clamp")

    def add(self, x_raw, group_ids=None):
        x_raw = _f32(x_raw)
        g = None if group_ids is None else np.ascontiguousarray(group_ids, dtype=np.int32)
        _check(load().b200nn_mpq_add(self.h, _vp(x_raw), C.c_size_t(x_raw.shape[0]), _vp(g)), "mpq_add")

    def search(self, q_raw, k, nprobe=1):
        q_raw = _f32(q_raw)
        nq = q_raw.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.uint64)
        _check(load().b200nn_mpq_search(self.h, _vp(q_raw), C.c_size_t(nq), C.c_int(nprobe), C.c_size_t(k), _vp(D), _vp(I)), "mpq_search")
        return D, I

    def search_host_ptr(self, q_ptr: int, nq: int, k: int, nprobe: int, out_dist_ptr: int, out_id_ptr: int):
        _check(load().b200nn_mpq_search(self.h, C.c_void_p(q_ptr), C.c_size_t(nq), C.c_int(nprobe), C.c_size_t(k), C.c_void_p(out_dist_ptr),
                                        C.c_void_p(out_id_ptr)), "mpq_search")

    def scores(self, q_raw, nprobe=3):
        q_raw = _f32(q_raw)
        out = np.empty((q_raw.shape[0], self.n_groups), dtype=np.float32)
        _check(load().b200nn_mpq_scores(self.h, _vp(q_raw), C.c_size_t(q_raw.shape[0]), C.c_int(nprobe), _vp(out)), "mpq_scores")
        return out

    def save_index(self, dir_or_path: str, group_paths=None):
        arr, n = None, 0
        if group_paths is not None:
            arr = (C.c_char_p * len(group_paths))(*[p.encode() for p in group_paths])
            n = len(group_paths)
        _check(load().b200nn_mpq_save_index(self.h, dir_or_path.encode(), arr, C.c_size_t(n)), "mpq_save_index")


class FlatIndex:
    """hnswlib::BruteforceSearch<float|int> drop-in surface (brute_force_search/src/brutoforce.hpp)."""
    METRIC = {"ip": 0, "l2": 1, "l2_u8": 2}

    def __init__(self, ctx: Context, metric: str, dim: int, max_elements: int, order: int = 4, _handle=None):
        self.ctx, self.metric, self.dim = ctx, self.METRIC[metric], dim
        self.h = C.c_void_p()
        if _handle is not None:
            self.h = _handle
        else:
            _check(load().b200nn_flat_create(ctx.h, C.c_int(self.metric), C.c_int(order), C.c_size_t(dim),
                                             C.c_size_t(max_elements), C.byref(self.h)), "flat_create")
        ctx._adopt(self)

    @classmethod
    def load_file(cls, ctx, metric: str, dim: int, path: str, order: int = 4):
        h = C.c_void_p()
        _check(load().b200nn_flat_load(ctx.h, C.c_int(cls.METRIC[metric]), C.c_int(order), C.c_size_t(dim), path.encode(),
                                       C.byref(h)), "flat_load")
        return cls(ctx, metric, dim, 0, order, _handle=h)

    @classmethod
    def load_hnsw_file(cls, ctx, metric: str, dim: int, path: str, order: int = 4):
        """vectors + labels of a HierarchicalNSW::saveIndex file as an exact flat index (hnswalg.h:491-519)."""
        h = C.c_void_p()
        _check(load().b200nn_flat_load_hnsw(ctx.h, C.c_int(cls.METRIC[metric]), C.c_int(order), C.c_size_t(dim), path.encode(),
                                            C.byref(h)), "flat_load_hnsw")
        return cls(ctx, metric, dim, 0, order, _handle=h)

    def info(self):
        mx, n = C.c_size_t(), C.c_size_t()
        _check(load().b200nn_flat_info(self.h, C.byref(mx), C.byref(n), None, C.c_size_t(0)), "flat_info")
        labels = np.zeros(n.value, dtype=np.uint64)
        _check(load().b200nn_flat_info(self.h, None, None, _vp(labels), C.c_size_t(labels.shape[0])), "flat_info")
        return int(mx.value), int(n.value), labels

    def close(self):
        if self.h:
            load().b200nn_flat_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _elems(self, a):
        return np.ascontiguousarray(a, dtype=np.uint8 if self.metric == 2 else np.float32)

    def add(self, vectors, labels):
        v = self._elems(vectors)
        l = np.ascontiguousarray(labels, dtype=np.uint64)
        _check(load().b200nn_flat_add(self.h, _vp(v), _vp(l), C.c_size_t(v.shape[0])), "flat_add")

    def remove(self, label: int):
        _check(load().b200nn_flat_remove(self.h, C.c_uint64(label)), "flat_remove")

    def __len__(self):
        v = C.c_size_t()
        _check(load().b200nn_flat_size(self.h, C.byref(v)), "flat_size")
        return int(v.value)

    def search(self, queries, k: int):
        q = self._elems(queries)
        nq = q.shape[0]
        D = np.empty((nq, k), dtype=np.int32 if self.metric == 2 else np.float32)
        L = np.empty((nq, k), dtype=np.uint64)
        _check(load().b200nn_flat_search(self.h, _vp(q), C.c_size_t(nq), C.c_size_t(k), _vp(D), _vp(L)), "flat_search")
        return D, L

    def search_dev(self, q_dev_ptr: int, nq: int, k: int, out_dist_ptr: int, out_label_ptr: int):
        _check(load().b200nn_flat_search_dev(self.h, C.c_void_p(q_dev_ptr), C.c_size_t(nq), C.c_size_t(k),
                                             C.c_void_p(out_dist_ptr), C.c_void_p(out_label_ptr)), "flat_search_dev")

    def save(self, path: str):
        _check(load().b200nn_flat_save(self.h, path.encode()), "flat_save")


class SQ:
    """cvtk::quant::Int8Quan drop-in surface (scalar_quantization/scalar_quantization/int8_quan.h)."""

    def __init__(self, ctx: Context, vmin, vdiff):
        self.ctx = ctx
        vmin, vdiff = _f32(vmin), _f32(vdiff)
        self.d = vmin.shape[0]
        self.h = C.c_void_p()
        _check(load().b200nn_sq_create(ctx.h, C.c_int(self.d), _vp(vmin), _vp(vdiff), C.byref(self.h)), "sq_create")
        ctx._adopt(self)

    def close(self):
        if self.h:
            load().b200nn_sq_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def train_minmax(ctx: Context, x_normed):
        x = _f32(x_normed)
        vmin = np.empty(x.shape[1], dtype=np.float32)
        vdiff = np.empty(x.shape[1], dtype=np.float32)
        _check(load().b200nn_sq_train_minmax(ctx.h, C.c_int(x.shape[1]), _vp(x), C.c_size_t(x.shape[0]), _vp(vmin), _vp(vdiff)),
               "sq_train_minmax")
        return vmin, vdiff

    def encode(self, x, l2norm=True):
        """Returns (codes, x_after): like the reference, x is normalised in place when l2norm."""
        x = _f32(x).copy()
        codes = np.empty(x.shape, dtype=np.uint8)
        _check(load().b200nn_sq_encode(self.h, _vp(x), C.c_size_t(x.shape[0]), C.c_int(int(l2norm)), _vp(codes)), "sq_encode")
        return codes, x

    def decode(self, codes, faiss_float=False):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        x = np.empty(codes.shape, dtype=np.float32)
        _check(load().b200nn_sq_decode(self.h, _vp(codes), C.c_size_t(codes.shape[0]), C.c_int(int(faiss_float)), _vp(x)), "sq_decode")
        return x


class Projection:
    """cvtk::PCAUtils drop-in surface (pca_train_project/pca_online/pca_utils.h): reduceDim = PCA project + L2 normalise."""

    def __init__(self, ctx: Context, vectors, mean=None):
        self.ctx = ctx
        vectors = _f32(vectors)
        self.N, self.K = vectors.shape
        mean_a = None if mean is None else _f32(mean).reshape(-1)
        self.h = C.c_void_p()
        _check(load().b200nn_proj_create(ctx.h, C.c_int(self.K), C.c_int(self.N), _vp(mean_a), _vp(vectors), C.byref(self.h)), "proj_create")
        ctx._adopt(self)

    @classmethod
    def load_model(cls, ctx: Context, path: str):
        """PCAUtils::loadModel: cv::PCA's YAML file (pca_train_project/model/*.yml)."""
        mean, vectors, _ = pca_read_model(path)
        return cls(ctx, vectors, mean)

    def close(self):
        if self.h:
            load().b200nn_proj_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reduce_dim(self, x, l2norm=True):
        x = _f32(x)
        y = np.empty((x.shape[0], self.N), dtype=np.float32)
        _check(load().b200nn_proj_apply(self.h, _vp(x), C.c_size_t(x.shape[0]), C.c_int(int(l2norm)), _vp(y)), "proj_apply")
        return y

    def reduce_dim_dev(self, x_dev_ptr: int, n: int, y_dev_ptr: int, l2norm=True):
        _check(load().b200nn_proj_apply_dev(self.h, C.c_void_p(x_dev_ptr), C.c_size_t(n), C.c_int(int(l2norm)), C.c_void_p(y_dev_ptr)),
               "proj_apply_dev")


def rootsift(ctx: Context, x, eps: float = 1e-7):
    """siftsIDX::rootSift on the device; returns a new array (the reference works in place)."""
    y = _f32(x).copy()
    _check(load().b200nn_rootsift(ctx.h, _vp(y), C.c_size_t(y.shape[0]), C.c_int(y.shape[1]), C.c_float(eps)), "rootsift")
    return y


def kmeans(ctx: Context, x, k: int, max_iter: int = 0, seed: int = 0):
    """Deterministic Lloyd k-means on the device (replaces yael's kmeans, train_PQ_codebook.cpp:164,229).
    Returns (centroids [k,d], assign [n], dist [n], iterations, mse)."""
    x = _f32(x)
    n, d = x.shape
    c = np.empty((k, d), dtype=np.float32)
    a = np.empty(n, dtype=np.int32)
    dist = np.empty(n, dtype=np.float32)
    it = C.c_int(0)
    mse = C.c_double(0.0)
    _check(load().b200nn_kmeans(ctx.h, _vp(x), C.c_size_t(n), C.c_int(d), C.c_int(k), C.c_int(max_iter), C.c_uint64(seed), _vp(c), _vp(a),
                                _vp(dist), C.byref(it), C.byref(mse)), "kmeans")
    return c, a, dist, it.value, mse.value


def pq_train(ctx: Context, x_raw, K: int, M: int, ksub: int = 256, perm=None, max_iter: int = 0, seed: int = 0):
    """TrainPQ::IFVPQ (CoarseQuan + ProdQuan) on the device.  K == 0 trains the flat-ADC model (one zero centroid).
    Returns (coarse [max(K,1),D], codebooks [M,ksub,D/M], mse [1+M])."""
    x_raw = _f32(x_raw)
    n, D = x_raw.shape
    perm_a = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    coarse = np.empty((max(K, 1), D), dtype=np.float32)
    cb = np.empty((M, ksub, D // max(M, 1)), dtype=np.float32)
    mse = np.zeros(1 + M, dtype=np.float64)
    _check(load().b200nn_pq_train(ctx.h, _vp(x_raw), C.c_size_t(n), C.c_int(D), C.c_int(K), C.c_int(M), C.c_int(ksub), _vp(perm_a),
                                  C.c_int(max_iter), C.c_uint64(seed), _vp(coarse), _vp(cb), _vp(mse)), "pq_train")
    return coarse, cb, mse


def pq_write_model(path: str, coarse, codebooks, perm=None):
    """TrainPQ::SaveCodebook's file (the one IVFOPQ::LoadModel reads)."""
    coarse, codebooks = _f32(coarse), _f32(codebooks)
    K, D = coarse.shape
    M, ksub, _ = codebooks.shape
    perm_a = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    _check(load().b200nn_pq_write_model(path.encode(), C.c_int(D), C.c_int(K), C.c_int(M), C.c_int(ksub), _vp(coarse), _vp(codebooks),
                                        _vp(perm_a)), "pq_write_model")


def kmeans_init_rows(n: int, k: int, seed: int):
    """host-only: rows the initial centroids are copied from."""
    r = np.empty(k, dtype=np.int32)
    _check(load().b200nn_kmeans_init_rows(C.c_size_t(n), C.c_int(k), C.c_uint64(seed), _vp(r)), "kmeans_init_rows")
    return r


def kmeans_plan_update(assign, dist, k: int):
    """host-only: one update's integer bookkeeping.  Returns (assign with donors moved, count, row_sorted, cluster_off)."""
    a = np.ascontiguousarray(assign, dtype=np.int32).copy()
    dist = _f32(dist)
    n = a.shape[0]
    count = np.empty(k, dtype=np.int32)
    rows = np.empty(n, dtype=np.int32)
    off = np.empty(k + 1, dtype=np.int64)
    _check(load().b200nn_kmeans_plan_update(C.c_size_t(n), C.c_int(k), _vp(a), _vp(dist), _vp(count), _vp(rows), _vp(off)), "kmeans_plan_update")
    return a, count, rows, off


def pca_read_model(path: str):
    """host-only: (mean [K], vectors [N,K], values [N]) of a cv::PCA YAML model (PCAUtils::loadModel, pca_utils.cc:16-23)."""
    K, N = C.c_int(0), C.c_int(0)
    _check(load().b200nn_pca_read_model(path.encode(), C.byref(K), C.byref(N), None, None, None), "pca_read_model")
    mean = np.empty(K.value, dtype=np.float32)
    vectors = np.empty((N.value, K.value), dtype=np.float32)
    values = np.empty(N.value, dtype=np.float32)
    _check(load().b200nn_pca_read_model(path.encode(), C.byref(K), C.byref(N), _vp(mean), _vp(vectors), _vp(values)), "pca_read_model")
    return mean, vectors, values
