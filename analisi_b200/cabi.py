"""ctypes binding of the C ABI in ``include/agofrt.h`` (tests and ``bench.py`` go through this).

Nothing here computes: every call lands in ``libagofrt.so``.  Loading fails loudly when the library
is missing, and every compute call fails with ``AgofrtError`` when no B200 is usable -- there is no
CPU path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGOFRT_LIB", os.path.join(_HERE, "libagofrt.so"))  # AGOFRT_LIB: tuning builds only

OK = 0
ERR_ARG, ERR_CUDA, ERR_WINDOW, ERR_NCCL, ERR_NONFINITE, ERR_TOO_LARGE, ERR_INTERNAL, ERR_RETYPED = -1, -2, -3, -4, -5, -6, -7, -8
OPT_EDGES, OPT_FORCE_GENERAL, OPT_NO_AGGREGATE, OPT_AGGREGATE, OPT_NO_SAFE, OPT_DENSE, OPT_SPARSE, OPT_NO_UBOX = 1, 2, 4, 8, 16, 32, 64, 128
OPT_NO_SMALL, OPT_ON_DEVICE, OPT_SMALL, OPT_SAFE2, OPT_SKEW, OPT_EXPLICIT_JOBS = 256, 512, 1024, 2048, 4096, 8192
SMALL_DEFAULT_SLOTS, SMALL_MAX_SLOTS = 256, 512   # kSmallDefault, kSmallMax of the library
MODE_BIT_SMALL = 1 << 8   # Stats.kernel_modes: the small-system kernel ran
COMM_ID_BYTES = 128
UP_WRAP, UP_WRITEBACK, UP_SHARED = 1, 2, 4

# every symbol include/agofrt.h declares (tests check the library exports all of them)
SYMBOLS = [
    "agofrt_version", "agofrt_last_error", "agofrt_device_count", "agofrt_host_alloc", "agofrt_host_free",
    "agofrt_ctx_create", "agofrt_ctx_destroy", "agofrt_ctx_ndev", "agofrt_comm_unique_id", "agofrt_comm_join",
    "agofrt_ctx_set_shard", "agofrt_shard_range", "agofrt_block_share", "agofrt_traj_create", "agofrt_traj_destroy", "agofrt_traj_upload", "agofrt_traj_upload_wrap", "agofrt_plan_retarget",
    "agofrt_traj_download_frame", "agofrt_pbc_wrap", "agofrt_traj_d2_all", "agofrt_traj_d2_pair", "agofrt_plan_create",
    "agofrt_plan_destroy", "agofrt_plan_thresholds", "agofrt_block", "agofrt_neighbour_hist", "agofrt_traj_set_cm", "agofrt_msd", "agofrt_fp64_peak",
    "agofrt_blockavg_create", "agofrt_blockavg_destroy", "agofrt_blockavg_begin", "agofrt_blockavg_push", "agofrt_blockavg_end",
    "agofrt_plan_last_counts", "agofrt_plan_info", "agofrt_traj_upload_ex", "agofrt_traj_download",
    "agofrt_blocks", "agofrt_plan_block_counts", "agofrt_blockavg_push_blocks",
    "agofrt_traj_set_ids", "agofrt_traj_upload_records", "agofrt_traj_set_rotation", "agofrt_traj_get_rotation",
    "agofrt_neighbours", "agofrt_sh_density",
]


class AgofrtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("agofrt error %d: %s" % (code, message))
        self.code = code


class Stats(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_double),
        ("total_ms", C.c_double),
        ("pair_evals", C.c_uint64),
        ("pair_evals_total", C.c_uint64),
        ("jobs", C.c_uint64),
        ("jobs_fast", C.c_uint64),
        ("launches", C.c_uint32),
        ("ndev_local", C.c_uint32),
        ("world", C.c_uint32),
        ("kernel_modes", C.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ }


_lib = None


def lib():
    """Load libagofrt.so (built in-tree by ``analisi_b200.build``).  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -m analisi_b200.build` (nvcc, sm_100a). "
            "analisi_b200 has no CPU or PyTorch fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    u64p = C.POINTER(C.c_uint64)
    L.agofrt_version.restype = C.c_char_p
    L.agofrt_last_error.restype = C.c_char_p
    L.agofrt_device_count.argtypes = [ip]
    L.agofrt_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.agofrt_host_free.argtypes = [vp]
    L.agofrt_ctx_create.argtypes = [C.POINTER(vp), ip, C.c_int]
    L.agofrt_ctx_destroy.argtypes = [vp]
    L.agofrt_ctx_ndev.argtypes = [vp]
    L.agofrt_comm_unique_id.argtypes = [C.c_char_p]
    L.agofrt_comm_join.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
    L.agofrt_ctx_set_shard.argtypes = [vp, C.c_int, C.c_int]
    L.agofrt_shard_range.argtypes = [C.c_uint64, C.c_int, C.c_int, u64p, u64p]
    L.agofrt_block_share.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    L.agofrt_traj_create.argtypes = [C.POINTER(vp), vp, C.c_size_t, C.c_int, ip, C.c_int, C.c_size_t]
    L.agofrt_traj_destroy.argtypes = [vp]
    L.agofrt_traj_upload.argtypes = [vp, C.c_size_t, C.c_size_t, vp, vp]
    L.agofrt_traj_upload_wrap.argtypes = [vp, C.c_size_t, C.c_size_t, vp, vp]
    L.agofrt_plan_retarget.argtypes = [vp, vp]
    L.agofrt_traj_upload_ex.argtypes = [vp, C.c_size_t, C.c_size_t, vp, vp, C.c_uint, vp]
    L.agofrt_traj_download.argtypes = [vp, C.c_size_t, C.c_size_t, vp]
    L.agofrt_traj_set_ids.argtypes = [vp, ip, ip]
    L.agofrt_neighbours.argtypes = [vp, C.c_size_t, u64p, dp, C.c_int, u64p, dp]
    L.agofrt_sh_density.argtypes = [vp, C.c_size_t, C.c_int, C.c_uint, dp, dp, ip]
    L.agofrt_traj_set_rotation.argtypes = [vp, C.c_size_t, C.c_size_t, dp]
    L.agofrt_traj_get_rotation.argtypes = [vp, C.c_size_t, dp]
    L.agofrt_traj_upload_records.argtypes = [vp, C.c_size_t, C.c_size_t, C.POINTER(vp), ip, C.POINTER(C.c_size_t), vp, C.c_uint, vp]
    L.agofrt_traj_download_frame.argtypes = [vp, C.c_size_t, dp]
    L.agofrt_pbc_wrap.argtypes = [vp, vp, C.c_size_t, C.c_size_t, dp, C.c_int]
    L.agofrt_traj_d2_all.argtypes = [vp, C.c_size_t, C.c_size_t, dp]
    L.agofrt_traj_d2_pair.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, dp]
    L.agofrt_plan_create.argtypes = [C.POINTER(vp), vp, C.c_double, C.c_double, C.c_uint]
    L.agofrt_plan_destroy.argtypes = [vp]
    L.agofrt_plan_thresholds.argtypes = [vp, dp]
    L.agofrt_plan_info.argtypes = [vp, ip, ip, ip]
    L.agofrt_block.argtypes = [vp, C.c_size_t, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, u64p, u64p,
                               C.POINTER(Stats)]
    L.agofrt_neighbour_hist.argtypes = [vp, C.c_double, C.c_size_t, C.c_uint, C.c_uint, u64p, C.POINTER(Stats)]
    L.agofrt_traj_set_cm.argtypes = [vp, C.c_size_t, C.c_size_t, dp]
    L.agofrt_msd.argtypes = [vp, C.c_size_t, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_int, dp, C.POINTER(Stats)]
    L.agofrt_fp64_peak.argtypes = [vp, C.c_int, C.c_double, dp]
    L.agofrt_blockavg_create.argtypes = [C.POINTER(vp), vp]
    L.agofrt_blockavg_destroy.argtypes = [vp]
    L.agofrt_blockavg_begin.argtypes = [vp, C.c_size_t]
    L.agofrt_blockavg_push.argtypes = [vp, vp, C.c_double]
    L.agofrt_blockavg_end.argtypes = [vp, C.c_uint, dp, dp]
    L.agofrt_plan_last_counts.argtypes = [vp, u64p, C.c_size_t]
    L.agofrt_blocks.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.POINTER(Stats)]
    L.agofrt_plan_block_counts.argtypes = [vp, C.c_uint, u64p, C.c_size_t]
    L.agofrt_blockavg_push_blocks.argtypes = [vp, vp, C.c_double]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("agofrt_version", "agofrt_last_error"):
            fn.restype = C.c_int
    _lib = L
    return L


def _check(rc):
    if rc != OK:
        raise AgofrtError(rc, lib().agofrt_last_error().decode("utf-8", "replace"))


def device_count():
    n = C.c_int(0)
    rc = lib().agofrt_device_count(C.byref(n))
    return n.value if rc == OK else 0


def shard_range(units, rank, world):
    b, e = C.c_uint64(0), C.c_uint64(0)
    _check(lib().agofrt_shard_range(int(units), int(rank), int(world), C.byref(b), C.byref(e)))
    return int(b.value), int(e.value)


def block_share(nblocks, rank, world, block):
    """(part_a, part_b): rank `rank` of `world` takes the work units [part_a, part_b) / world of `block` (agofrt_blocks)."""
    a, b = C.c_uint(0), C.c_uint(0)
    _check(lib().agofrt_block_share(int(nblocks), int(rank), int(world), int(block), C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class PinnedArray:
    """A float64 numpy array in page-locked host memory (agofrt_host_alloc)."""

    def __init__(self, shape):
        n = int(np.prod(shape))
        self._ptr = C.c_void_p()
        _check(lib().agofrt_host_alloc(C.byref(self._ptr), n * 8))
        buf = (C.c_double * n).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)

    def free(self):
        if self._ptr:
            self.array = None
            lib().agofrt_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    def __init__(self, devices=None):
        import weakref
        self._children = weakref.WeakSet()   # windows and plans: they must be destroyed before the context
        self._h = C.c_void_p()
        if devices is None:
            rc = lib().agofrt_ctx_create(C.byref(self._h), None, 0)
        elif devices == "all":
            rc = lib().agofrt_ctx_create(C.byref(self._h), None, -1)
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = lib().agofrt_ctx_create(C.byref(self._h), arr, len(devices))
        _check(rc)

    @property
    def ndev(self):
        return lib().agofrt_ctx_ndev(self._h)

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(COMM_ID_BYTES)
        _check(lib().agofrt_comm_unique_id(buf))
        return buf.raw

    def join(self, comm_id, first_rank, world):
        assert len(comm_id) == COMM_ID_BYTES
        _check(lib().agofrt_comm_join(self._h, comm_id, first_rank, world))

    def set_shard(self, first_rank, world):
        _check(lib().agofrt_ctx_set_shard(self._h, first_rank, world))

    def pbc_wrap(self, pos, box_internal):
        """In-place BaseTrajectory::pbc_wrap of pos[F,N,3] (float64, C-contiguous)."""
        assert pos.dtype == np.float64 and pos.flags.c_contiguous and pos.ndim == 3
        box = np.ascontiguousarray(box_internal, dtype=np.float64)
        assert box.shape[0] == pos.shape[0] and box.shape[1] in (6, 9)
        _check(lib().agofrt_pbc_wrap(self._h, pos.ctypes.data, pos.shape[0], pos.shape[1], _dp(box), box.shape[1]))
        return pos

    def fp64_peak(self, seconds=0.5, local_device=0):
        out = C.c_double(0)
        _check(lib().agofrt_fp64_peak(self._h, local_device, seconds, C.byref(out)))
        return out.value

    def close(self):
        if self._h:
            for child in list(self._children):
                child.close()
            lib().agofrt_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceTrajectory:
    """Device-resident window: positions [frame][atom][3], internal box rows, dense type ids."""

    def __init__(self, ctx, natoms, box_stride, type_id, ntypes, max_frames):
        self.ctx = ctx
        self.natoms, self.box_stride, self.ntypes = int(natoms), int(box_stride), int(ntypes)
        tid = np.ascontiguousarray(type_id, dtype=np.int32)
        assert tid.shape == (self.natoms,)
        self._h = C.c_void_p()
        _check(lib().agofrt_traj_create(C.byref(self._h), ctx._h, self.natoms, self.box_stride,
                                        tid.ctypes.data_as(C.POINTER(C.c_int)), self.ntypes, int(max_frames)))
        ctx._children.add(self)

    def upload(self, first_frame, pos, box_internal):
        assert pos.dtype == np.float64 and pos.flags.c_contiguous
        assert pos.ndim == 3 and pos.shape[1] == self.natoms and pos.shape[2] == 3
        box = np.ascontiguousarray(box_internal, dtype=np.float64)
        assert box.shape == (pos.shape[0], self.box_stride)
        _check(lib().agofrt_traj_upload(self._h, int(first_frame), pos.shape[0], pos.ctypes.data, box.ctypes.data))

    def upload_wrap(self, first_frame, pos, box_internal):
        """Upload with BaseTrajectory::pbc_wrap applied on the device; ``pos`` is wrapped in place."""
        assert pos.dtype == np.float64 and pos.flags.c_contiguous and pos.flags.writeable
        assert pos.ndim == 3 and pos.shape[1] == self.natoms and pos.shape[2] == 3
        box = np.ascontiguousarray(box_internal, dtype=np.float64)
        assert box.shape == (pos.shape[0], self.box_stride)
        _check(lib().agofrt_traj_upload_wrap(self._h, int(first_frame), pos.shape[0], pos.ctypes.data, box.ctypes.data))

    def upload_ex(self, first_frame, pos, box_internal, wrap=False, shared=False, out=None):
        """agofrt_traj_upload_ex: ``pos`` is only read (it may be pageable memory); with ``wrap`` the device wraps the
        frames, and ``out`` (an array like pos, or pos itself) receives the wrapped frames; ``shared`` deals the frames to
        the devices of the communicator and exchanges the shares device to device."""
        assert pos.dtype == np.float64 and pos.flags.c_contiguous
        assert pos.ndim == 3 and pos.shape[1] == self.natoms and pos.shape[2] == 3
        box = np.ascontiguousarray(box_internal, dtype=np.float64)
        assert box.shape == (pos.shape[0], self.box_stride)
        flags = (UP_WRAP if wrap else 0) | (UP_SHARED if shared else 0) | (UP_WRITEBACK if out is not None else 0)
        if out is not None:
            assert wrap and out.dtype == np.float64 and out.flags.c_contiguous and out.shape == pos.shape and out.flags.writeable
        _check(lib().agofrt_traj_upload_ex(self._h, int(first_frame), pos.shape[0], pos.ctypes.data, box.ctypes.data, flags,
                                           out.ctypes.data if out is not None else None))

    def set_ids(self, slot_to_id, slot_raw_type):
        a = np.ascontiguousarray(slot_to_id, dtype=np.int32)
        b = np.ascontiguousarray(slot_raw_type, dtype=np.int32)
        assert a.shape == b.shape == (self.natoms,)
        _check(lib().agofrt_traj_set_ids(self._h, a.ctypes.data_as(C.POINTER(C.c_int)), b.ctypes.data_as(C.POINTER(C.c_int))))

    def upload_records(self, first_frame, frames, box_internal, wrap=False, shared=False, out=None):
        """agofrt_traj_upload_records: ``frames`` is a list (one entry per frame) of lists of float64 arrays [n][8]
        (the chunks of the frame: id type x y z vx vy vz per atom)."""
        chunks = [np.ascontiguousarray(c, dtype=np.float64) for fr in frames for c in fr]
        assert all(c.ndim == 2 and c.shape[1] == 8 for c in chunks)
        ptrs = (C.c_void_p * max(len(chunks), 1))(*[c.ctypes.data for c in chunks])
        atoms = (C.c_int * max(len(chunks), 1))(*[c.shape[0] for c in chunks])
        begin = np.cumsum([0] + [len(fr) for fr in frames])
        fc = (C.c_size_t * len(begin))(*[int(x) for x in begin])
        box = np.ascontiguousarray(box_internal, dtype=np.float64)
        assert box.shape == (len(frames), self.box_stride)
        flags = (UP_WRAP if wrap else 0) | (UP_SHARED if shared else 0) | (UP_WRITEBACK if out is not None else 0)
        _check(lib().agofrt_traj_upload_records(self._h, int(first_frame), len(frames), ptrs, atoms, fc, box.ctypes.data, flags,
                                                out.ctypes.data if out is not None else None))

    def neighbours(self, frame, spec, sort=False):
        """Neighbours::update_neigh: ``spec`` = [(max neighbours, cutoff^2), ...] per type.  Returns (counts [N][T],
        indices [N][T][maxn] (-1 past the count), r [N][T][maxn][4]) unpacked from the reference's own layout."""
        nn = np.array([s[0] for s in spec], dtype=np.uint64)
        c2 = np.array([s[1] for s in spec], dtype=np.float64)
        assert len(spec) == self.ntypes
        n = self.natoms
        words = int(((nn + 1) * n).sum())
        doubles = int((nn * n * 4).sum())
        lst = np.zeros(words, dtype=np.uint64)
        rpos = np.zeros(max(doubles, 1), dtype=np.float64)
        _check(lib().agofrt_neighbours(self._h, int(frame), nn.ctypes.data_as(C.POINTER(C.c_uint64)), _dp(c2), int(bool(sort)),
                                       lst.ctypes.data_as(C.POINTER(C.c_uint64)), _dp(rpos)))
        maxn = int(nn.max())
        counts = np.zeros((n, self.ntypes), dtype=np.int64)
        idx = -np.ones((n, self.ntypes, maxn), dtype=np.int64)
        r = np.zeros((n, self.ntypes, maxn, 4))
        lo = ro = 0
        for t in range(self.ntypes):
            k = int(nn[t])
            blk = lst[lo:lo + (k + 1) * n].reshape(n, k + 1)
            rb = rpos[ro:ro + k * n * 4].reshape(n, k, 4)
            counts[:, t] = blk[:, 0]
            for i in range(n):
                c = int(blk[i, 0])
                idx[i, t, :c] = blk[i, 1:1 + c]
                r[i, t, :c] = rb[i, :c]
            lo += (k + 1) * n
            ro += k * n * 4
        return counts, idx, r

    def sh_density(self, frame, lmax, nbin, rminmax):
        """SphericalBase::calc: (result [N][T][nbin][(lmax+1)^2], counter [N][T][nbin])."""
        rm = np.ascontiguousarray(rminmax, dtype=np.float64).reshape(self.ntypes * self.ntypes, 2)
        res = np.zeros((self.natoms, self.ntypes, int(nbin), (lmax + 1) ** 2), dtype=np.float64)
        cnt = np.zeros((self.natoms, self.ntypes, int(nbin)), dtype=np.int32)
        _check(lib().agofrt_sh_density(self._h, int(frame), int(lmax), int(nbin), _dp(rm), _dp(res), cnt.ctypes.data_as(C.POINTER(C.c_int))))
        return res, cnt

    def set_rotation(self, first_frame, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, 9)
        _check(lib().agofrt_traj_set_rotation(self._h, int(first_frame), q.shape[0], _dp(q)))

    def get_rotation(self, frame):
        out = np.zeros(9, dtype=np.float64)
        _check(lib().agofrt_traj_get_rotation(self._h, int(frame), _dp(out)))
        return out

    def download(self, first_frame, nframes):
        """Frames of the device window in the caller's atom order (wrapped if uploaded with wrap)."""
        out = np.empty((int(nframes), self.natoms, 3), dtype=np.float64)
        _check(lib().agofrt_traj_download(self._h, int(first_frame), int(nframes), out.ctypes.data))
        return out

    def download_frame(self, frame):
        out = np.empty((self.natoms, 3), dtype=np.float64)
        _check(lib().agofrt_traj_download_frame(self._h, int(frame), _dp(out)))
        return out

    def d2_all(self, frame_i, frame_j):
        out = np.zeros((self.natoms, self.natoms, 4), dtype=np.float64)
        _check(lib().agofrt_traj_d2_all(self._h, int(frame_i), int(frame_j), _dp(out)))
        return out

    def neighbour_hist(self, r, tstart, ntimesteps, skip=1, hist=None):
        """IstogrammaAtomiRaggio::calculate on the uploaded window; returns (hist[ntypes][natoms+1], stats)."""
        if hist is None:
            hist = np.zeros((self.ntypes, self.natoms + 1), dtype=np.uint64)
        st = Stats()
        _check(lib().agofrt_neighbour_hist(self._h, float(r), int(tstart), int(ntimesteps), int(skip),
                                           hist.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(st)))
        return hist, st.as_dict()

    def set_cm(self, first_frame, cm):
        cm = np.ascontiguousarray(cm, dtype=np.float64)
        assert cm.ndim == 3 and cm.shape[1:] == (self.ntypes, 3)
        _check(lib().agofrt_traj_set_cm(self._h, int(first_frame), cm.shape[0], _dp(cm)))

    def msd(self, primo, ntimesteps, lmax=0, skip=1, cm_msd=False, cm_self=False):
        """MSD<T>::calculate(primo) after reset(ntimesteps); returns (vdata [leff][f_cm][ntypes], stats)."""
        leff = gofrt_leff(int(ntimesteps), int(lmax))
        out = np.zeros((leff, 2 if cm_msd else 1, self.ntypes), dtype=np.float64)
        st = Stats()
        _check(lib().agofrt_msd(self._h, int(primo), int(ntimesteps), leff, int(skip), int(bool(cm_msd)), int(bool(cm_self)),
                                _dp(out), C.byref(st)))
        return out, st.as_dict()

    def d2_pair(self, i, j, frame_i, frame_j):
        out = np.zeros(4, dtype=np.float64)
        _check(lib().agofrt_traj_d2_pair(self._h, int(i), int(j), int(frame_i), int(frame_j), _dp(out)))
        return out

    def close(self):
        if self._h:
            lib().agofrt_traj_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Plan:
    def __init__(self, traj, rmin, rmax, nbin):
        self.traj = traj
        self.nbin = int(nbin)
        self._h = C.c_void_p()
        _check(lib().agofrt_plan_create(C.byref(self._h), traj._h, float(rmin), float(rmax), self.nbin))
        traj.ctx._children.add(self)

    def retarget(self, traj):
        _check(lib().agofrt_plan_retarget(self._h, traj._h))
        self.traj = traj

    def info(self):
        """{'safe_zone': bool, 'two_floor': bool, 'guard_bins': int}: which float shortcuts passed their validation."""
        a, b, g = C.c_int(0), C.c_int(0), C.c_int(0)
        _check(lib().agofrt_plan_info(self._h, C.byref(a), C.byref(b), C.byref(g)))
        return {"safe_zone": bool(a.value), "two_floor": bool(b.value), "guard_bins": int(g.value)}

    def thresholds(self):
        out = np.empty(self.nbin + 1, dtype=np.float64)
        _check(lib().agofrt_plan_thresholds(self._h, _dp(out)))
        return out

    def block(self, primo, ntimesteps, leff, skip=1, every=1, options=0, edges=False):
        """Integer counts [leff][ntypes*(ntypes+1)][nbin] of one calculate(primo) after reset(ntimesteps).

        Returns (counts, stats dict[, edge_pairs]).  With OPT_ON_DEVICE the counts stay on the device (for
        BlockAverage.push / last_counts) and ``counts`` is None."""
        nt = self.traj.ntypes
        on_device = bool(int(options) & OPT_ON_DEVICE)
        counts = None if on_device else np.zeros((int(leff), nt * (nt + 1), self.nbin), dtype=np.uint64)
        st = Stats()
        e = C.c_uint64(0)
        rc = lib().agofrt_block(self._h, int(primo), int(ntimesteps), int(leff), int(skip), int(every),
                                int(options) | (OPT_EDGES if edges else 0),
                                None if on_device else counts.ctypes.data_as(C.POINTER(C.c_uint64)),
                                C.byref(e) if edges else None, C.byref(st))
        _check(rc)
        if edges:
            return counts, st.as_dict(), int(e.value)
        return counts, st.as_dict()

    def blocks(self, primo0, stride, nblocks, ntimesteps, leff, skip=1, every=1, options=0):
        """agofrt_blocks: nblocks whole blocks dealt to the devices; the counts stay on the devices.  Returns stats."""
        st = Stats()
        _check(lib().agofrt_blocks(self._h, int(primo0), int(stride), int(nblocks), int(ntimesteps), int(leff), int(skip),
                                   int(every), int(options), C.byref(st)))
        return st.as_dict()

    def block_counts(self, block, leff):
        nt = self.traj.ntypes
        counts = np.zeros((int(leff), nt * (nt + 1), self.nbin), dtype=np.uint64)
        _check(lib().agofrt_plan_block_counts(self._h, int(block), counts.ctypes.data_as(C.POINTER(C.c_uint64)), counts.size))
        return counts

    def last_counts(self, leff):
        """The counts of the last block() from the device (what a block without OPT_ON_DEVICE returns)."""
        nt = self.traj.ntypes
        counts = np.zeros((int(leff), nt * (nt + 1), self.nbin), dtype=np.uint64)
        _check(lib().agofrt_plan_last_counts(self._h, counts.ctypes.data_as(C.POINTER(C.c_uint64)), counts.size))
        return counts

    def close(self):
        if self._h:
            lib().agofrt_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BlockAverage:
    """MediaVar over blocks on the device (agofrt_blockavg_*): mean and variance of the mean of count*incr."""

    def __init__(self, ctx):
        self._h = C.c_void_p()
        _check(lib().agofrt_blockavg_create(C.byref(self._h), ctx._h))
        ctx._children.add(self)
        self.len = 0

    def begin(self, length):
        self.len = int(length)
        _check(lib().agofrt_blockavg_begin(self._h, self.len))

    def push(self, plan, incr):
        _check(lib().agofrt_blockavg_push(self._h, plan._h, float(incr)))

    def push_blocks(self, plan, incr):
        _check(lib().agofrt_blockavg_push_blocks(self._h, plan._h, float(incr)))

    def end(self, n_b):
        mean = np.empty(self.len, dtype=np.float64)
        var = np.empty(self.len, dtype=np.float64)
        _check(lib().agofrt_blockavg_end(self._h, int(n_b), _dp(mean), _dp(var)))
        return mean, var

    def close(self):
        if self._h:
            lib().agofrt_blockavg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- the host-side arithmetic of Gofrt that is not pair work (mirrors lib/src/gofrt.cpp) -------
def device_slots(type_id, pad=8):
    """Slots of the device layout: every type group padded to a multiple of 8 (kPadGroup)."""
    counts = np.bincount(np.asarray(type_id, dtype=np.int64))
    return int(((counts + pad - 1) // pad * pad).sum())


def gofrt_leff(ntimesteps, lmax):
    """Gofrt::reset, reference lib/src/gofrt.cpp:55."""
    return ntimesteps if (ntimesteps < lmax or lmax == 0) else lmax


def gofrt_incr(ntimesteps, skip):
    """Gofrt::calc_init, reference lib/src/gofrt.cpp:91-92."""
    skip = skip or 1
    q = ntimesteps // skip
    return 1.0 / int(q) if q > 0 else 1.0


def gofrt_nextra(total_frames, n_b, lmax):
    """Gofrt::nExtraTimesteps, reference lib/src/gofrt.cpp:37-39."""
    a = total_frames // (n_b + 1) + 1
    return a if (a < lmax or lmax == 0) else lmax
