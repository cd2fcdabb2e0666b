"""Host-side engine: one context (one GPU, one stream) over the C ABI.

Arrays are numpy `uint64` with a trailing limb axis -- field elements (n, 4), ExtendedPoint
(n, 20), AffinePoint (n, 8), ExtendedNielsPoint (n, 16), AffineNielsPoint (n, 12) -- or
`uint8` (n, 32) byte strings, i.e. exactly the reference's in-memory layouts
(src/lib.rs:81-84, 139-145, 255-259, 327-332; src/fr.rs:23).  `DeviceArray` keeps a batch
resident in HBM between calls.
"""
import ctypes as C

import numpy as np

from . import _lib as L

FQ, FR = "fq", "fr"
EXT_W, AFF_W, NIELS_W, ANIELS_W = 20, 8, 16, 12


class JubjubError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{L.ERRORS.get(code, code)}: {msg}")
        self.code = code


class DeviceArray:
    """A caller-owned buffer in the context's HBM (jj_malloc); shape/dtype mirror numpy."""

    def __init__(self, engine, shape, dtype):
        self.engine, self.shape, self.dtype = engine, tuple(shape), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        engine._check(engine.lib.jj_malloc(engine.ctx, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def __len__(self):
        return self.shape[0]

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.shape == self.shape, (host.shape, self.shape)
        self.engine._check(self.engine.lib.jj_memcpy_h2d(self.engine.ctx, self.ptr, host.ctypes.data, self.nbytes))
        return self

    def download(self):
        out = np.empty(self.shape, dtype=self.dtype)
        self.engine._check(self.engine.lib.jj_memcpy_d2h(self.engine.ctx, out.ctypes.data, self.ptr, self.nbytes))
        return out

    def free(self):
        if self.ptr:
            self.engine.lib.jj_free(self.engine.ctx, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    def __init__(self, device=0):
        self.lib = L.load()
        ctx = C.c_void_p()
        rc = self.lib.jj_init(device, C.byref(ctx))
        if rc != 0:
            raise JubjubError(rc, f"jj_init(device={device}) failed -- a B200 (sm_100) device is required")
        self.ctx, self.device = ctx, device
        self.nranks, self.rank = 1, 0

    def close(self):
        if self.ctx:
            self.lib.jj_destroy(self.ctx)
            self.ctx = None

    def _check(self, rc):
        if rc != 0:
            raise JubjubError(rc, self.lib.jj_last_error(self.ctx).decode())

    # ---- plumbing ---------------------------------------------------------------------------
    def empty(self, shape, dtype=np.uint64):
        return DeviceArray(self, shape, dtype)

    def pinned_empty(self, shape, dtype=np.uint64):
        """A numpy array in page-locked host memory (jj_host_alloc), freed when the array is garbage-collected.  Host
        batches passed in such arrays are copied by DMA without the driver's pageable staging, and the wire-format
        path (scalar_mul_encoded_vartime with out= / ok=... through the C ABI) reads and writes them in place."""
        import weakref

        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._check(self.lib.jj_host_alloc(self.ctx, nbytes, C.byref(p)))
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        eng_ref, addr = weakref.ref(self), p.value

        def _free():  # only through a context that is still alive (a closed engine's pinned buffers are left to the driver)
            e = eng_ref()
            if e is not None and e.ctx:
                e.lib.jj_host_free(e.ctx, C.c_void_p(addr))

        weakref.finalize(buf, _free)
        return arr

    def to_device(self, host):
        host = np.ascontiguousarray(host)
        return DeviceArray(self, host.shape, host.dtype).upload(host)

    def _arg(self, a, width, dtype=np.uint64):
        """-> (pointer, n, is_device, keepalive)"""
        if isinstance(a, DeviceArray):
            assert a.shape[1] == width and a.dtype == np.dtype(dtype), (a.shape, a.dtype, width)
            return a.ptr, a.shape[0], True, a
        a = np.ascontiguousarray(a, dtype=dtype)
        if a.ndim != 2 or a.shape[1] != width:
            raise ValueError(f"expected shape (n, {width}) {np.dtype(dtype)}, got {a.shape}")
        return a.ctypes.data, a.shape[0], False, a

    def _out(self, like_device, n, width, dtype=np.uint64, out=None):
        if out is not None:
            if isinstance(out, DeviceArray) != like_device:
                raise ValueError("inputs and output must all be host arrays or all DeviceArrays")
            # the kernel (or the D2H copy) writes n x width elements: a smaller or differently typed buffer is overrun
            if out.dtype != np.dtype(dtype) or len(out.shape) != 2 or out.shape[1] != width or out.shape[0] < n:
                raise ValueError(f"out must be ({n}, {width}) {np.dtype(dtype)} (more rows allowed), got {out.shape} {out.dtype}")
            if not like_device and not out.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be C-contiguous")
            return out
        return DeviceArray(self, (n, width), dtype) if like_device else np.empty((n, width), dtype=dtype)

    @staticmethod
    def _ptr(x):
        return x.ptr if isinstance(x, DeviceArray) else x.ctypes.data

    def _call(self, name, ins, out_width, out_dtype=np.uint64, flags=0, out=None, ok=False, extra_pre=(), ok_out=None):
        """ins: list of (array, width, dtype).  Returns out (and ok flags when ok=True)."""
        args, n, dev, keep = [], None, None, []
        for a, w, dt in ins:
            p, cnt, isdev, k = self._arg(a, w, dt)
            if n is None:
                n, dev = cnt, isdev
            elif cnt != n:
                # the reference panics on a length mismatch (assert_eq!, src/lib.rs:841)
                raise JubjubError(-1, f"length mismatch: {cnt} != {n}")
            elif isdev != dev:
                raise ValueError("inputs must all be host arrays or all DeviceArrays")
            args.append(p)
            keep.append(k)
        o = self._out(dev, n, out_width, out_dtype, out)
        call = list(extra_pre) + args + [self._ptr(o)]
        okbuf = None
        if ok:
            okbuf = self._out(dev, n, 1, np.uint8, ok_out if ok_out is None or isinstance(ok_out, DeviceArray) else ok_out.reshape(-1, 1))
            call.append(self._ptr(okbuf))
        f = flags | (L.JJ_DEVICE_PTRS if dev else 0)
        self._check(getattr(self.lib, name)(self.ctx, *call, n, f))
        if ok:
            return o, (okbuf if dev else okbuf.reshape(-1))
        return o

    # ---- field batches ---------------------------------------------------------------------
    def fe_mul(self, field, a, b, flags=0, out=None):
        return self._call(f"jj_{field}_mul", [(a, 4, np.uint64), (b, 4, np.uint64)], 4, flags=flags, out=out)

    def fe_add(self, field, a, b, flags=0, out=None):
        return self._call(f"jj_{field}_add", [(a, 4, np.uint64), (b, 4, np.uint64)], 4, flags=flags, out=out)

    def fe_sub(self, field, a, b, flags=0, out=None):
        return self._call(f"jj_{field}_sub", [(a, 4, np.uint64), (b, 4, np.uint64)], 4, flags=flags, out=out)

    def fe_square(self, field, a, flags=0, out=None):
        return self._call(f"jj_{field}_square", [(a, 4, np.uint64)], 4, flags=flags, out=out)

    def fe_neg(self, field, a, flags=0, out=None):
        return self._call(f"jj_{field}_neg", [(a, 4, np.uint64)], 4, flags=flags, out=out)

    def fe_double(self, field, a, flags=0, out=None):
        return self._call(f"jj_{field}_double", [(a, 4, np.uint64)], 4, flags=flags, out=out)

    def fe_invert(self, field, a, flags=0):
        """-> (inverse, ok): ok[i] = 0 and inverse[i] = 0 where a[i] = 0 (CtOption::none)."""
        return self._call(f"jj_{field}_invert", [(a, 4, np.uint64)], 4, flags=flags, ok=True)

    def fe_sqrt(self, field, a):
        """-> (root, ok): ok[i] = 0 and root[i] = 0 for a non-residue (src/fr.rs:384-399)."""
        return self._call(f"jj_{field}_sqrt", [(a, 4, np.uint64)], 4, ok=True)

    def fe_to_bytes(self, field, a):
        return self._call(f"jj_{field}_to_bytes", [(a, 4, np.uint64)], 32, np.uint8)

    def fe_from_bytes(self, field, b):
        return self._call(f"jj_{field}_from_bytes", [(b, 32, np.uint8)], 4, ok=True)

    def fe_from_bytes_wide(self, field, b64):
        return self._call(f"jj_{field}_from_bytes_wide", [(b64, 64, np.uint8)], 4)

    def fe_stream(self, field, seed, n, first=0, device=False):
        o = self._out(device, n, 4)
        f = L.JJ_DEVICE_PTRS if device else 0
        self._check(getattr(self.lib, f"jj_{field}_stream")(self.ctx, seed, first, self._ptr(o), n, f))
        return o

    # ---- points ----------------------------------------------------------------------------
    def point_double(self, p, out=None):
        return self._call("jj_point_double", [(p, EXT_W, np.uint64)], EXT_W, out=out)

    def point_add(self, p, q, subtract=False, out=None):
        return self._call("jj_point_add", [(p, EXT_W, np.uint64), (q, EXT_W, np.uint64)], EXT_W,
                          flags=L.JJ_SUBTRACT if subtract else 0, out=out)

    def point_add_niels(self, p, q, subtract=False, out=None):
        return self._call("jj_point_add_niels", [(p, EXT_W, np.uint64), (q, NIELS_W, np.uint64)], EXT_W,
                          flags=L.JJ_SUBTRACT if subtract else 0, out=out)

    def point_add_affine_niels(self, p, q, subtract=False, out=None):
        return self._call("jj_point_add_affine_niels", [(p, EXT_W, np.uint64), (q, ANIELS_W, np.uint64)], EXT_W,
                          flags=L.JJ_SUBTRACT if subtract else 0, out=out)

    def point_to_niels(self, p):
        return self._call("jj_point_to_niels", [(p, EXT_W, np.uint64)], NIELS_W)

    def affine_to_niels(self, p):
        return self._call("jj_affine_to_niels", [(p, AFF_W, np.uint64)], ANIELS_W)

    @staticmethod
    def _out_fmt(output):
        return {"extended": (EXT_W, np.uint64, 0), "affine": (AFF_W, np.uint64, L.JJ_OUT_AFFINE),
                "bytes": (32, np.uint8, L.JJ_OUT_BYTES)}[output]

    # The fast scalar-multiplication entry points are variable-time in the scalar (zero window digits skip their addition,
    # the table is indexed by the digit): the reference's `&ExtendedPoint * &Fr` is constant-time by policy
    # (src/lib.rs:12-17) and names every variable-time routine `*_vartime` (:14-15) -- so do these.  Public scalars only.
    # `scalar_mul` (no suffix) is the constant-time-in-the-scalar mode of the same kernel (JJ_CONST_TIME).
    def scalar_mul(self, points, scalars, output="extended", scalar_mont=False, out=None, flags=0):
        """out[i] = [scalars[i]] points[i] with no branch or memory address depending on the scalars (window table
        scanned, sign by selects, every addition executed -- the batch analogue of the reference's constant-time
        `&ExtendedPoint * &Fr`, src/lib.rs:356-379, 873-879).  Same results as scalar_mul_vartime, slower."""
        return self.scalar_mul_vartime(points, scalars, output=output, scalar_mont=scalar_mont, out=out,
                                       flags=flags | L.JJ_CONST_TIME)

    def scalar_mul_vartime(self, points, scalars, output="extended", scalar_mont=False, out=None, flags=0):
        """out[i] = [scalars[i]] points[i]  (`&ExtendedPoint * &Fr`, src/lib.rs:873-879); variable-time."""
        w, dt, f = self._out_fmt(output)
        sc = (scalars, 4, np.uint64) if scalar_mont else (scalars, 32, np.uint8)
        return self._call("jj_scalar_mul", [(points, EXT_W, np.uint64), sc], w, dt,
                          flags=f | flags | (L.JJ_SCALAR_MONT if scalar_mont else 0), out=out)

    def scalar_mul_encoded_vartime(self, encodings, scalars, output="bytes", zip216=True, check_subgroup=False, flags=0,
                                   out=None, ok_out=None):
        """(out, ok): out[i] = [scalars[i]] AffinePoint::from_bytes(encodings[i]) -- wire format in (32-byte
        encodings, src/lib.rs:455-464), decoded on the device (src/lib.rs:541-627); ok[i] = 0 for a rejected
        encoding (its output unit is unspecified).  check_subgroup: the decode is SubgroupPoint::from_bytes
        (src/lib.rs:1427-1429), i.e. ok[i] also requires is_torsion_free.  Variable-time.  With host arrays from
        pinned_empty() (inputs, out=, ok_out=) no staging copy is made: the kernels read and write them in place."""
        w, dt, f = self._out_fmt(output)
        f |= (0 if zip216 else L.JJ_PRE_ZIP216) | (L.JJ_CHECK_SUBGROUP if check_subgroup else 0) | flags
        return self._call("jj_scalar_mul_encoded", [(encodings, 32, np.uint8), (scalars, 32, np.uint8)], w, dt,
                          flags=f, ok=True, out=out, ok_out=ok_out)

    def scalar_mul_fixed_vartime(self, base_affine, scalars, output="extended", scalar_mont=False, out=None):
        """out[i] = [scalars[i]] base  (`&AffinePoint * &Fr`, src/lib.rs:1109-1115), one shared base; variable-time."""
        w, dt, f = self._out_fmt(output)
        base = np.ascontiguousarray(base_affine, dtype=np.uint64).reshape(1, AFF_W)
        if scalar_mont:
            p, n, dev, keep = self._arg(scalars, 4, np.uint64)
        else:
            p, n, dev, keep = self._arg(scalars, 32, np.uint8)
        o = self._out(dev, n, w, dt, out)
        bptr = base.ctypes.data
        bdev = None
        if dev:
            bdev = self.to_device(base)
            bptr = bdev.ptr
        fl = f | (L.JJ_DEVICE_PTRS if dev else 0) | (L.JJ_SCALAR_MONT if scalar_mont else 0)
        self._check(self.lib.jj_scalar_mul_fixed(self.ctx, bptr, p, self._ptr(o), n, fl))
        return o

    def batch_normalize(self, p, out=None):
        """ExtendedPoint::batch_normalize (src/lib.rs:840-858): ExtendedPoint -> AffinePoint."""
        return self._call("jj_batch_normalize", [(p, EXT_W, np.uint64)], AFF_W, out=out)

    def batch_normalize_to_bytes(self, p, out=None):
        """GroupEncoding::to_bytes for ExtendedPoint (src/lib.rs:1419-1421): normalise + encode in one pass."""
        return self._call("jj_batch_normalize", [(p, EXT_W, np.uint64)], 32, np.uint8, flags=L.JJ_OUT_BYTES, out=out)

    def batch_normalize_extended(self, p, in_place=False):
        """batch_normalize (src/lib.rs:1084-1107): the points themselves become (u/z, v/z, 1, u/z, v/z); in_place
        overwrites `p` like the reference's `&mut [ExtendedPoint]`."""
        return self._call("jj_batch_normalize_extended", [(p, EXT_W, np.uint64)], EXT_W, out=p if in_place else None)

    def point_sum(self, p, group_size=None, output="extended"):
        """Sum<ExtendedPoint> (src/lib.rs:183-193) over consecutive groups of `group_size` points (default: the whole
        batch -> one point).  Returns one point per group in the requested format."""
        w, dt, f = self._out_fmt(output)
        ptr, n, dev, keep = self._arg(p, EXT_W, np.uint64)
        g = n if group_size is None else int(group_size)
        if g < 0 or (g and n % g):
            raise JubjubError(-1, f"{n} points do not split into groups of {g}")
        groups = (n // g) if g else 1
        if n == 0:
            groups, g = 1, 0
        o = self._out(dev, groups, w, dt)
        self._check(self.lib.jj_point_sum(self.ctx, ptr, self._ptr(o), groups, g, f | (L.JJ_DEVICE_PTRS if dev else 0)))
        return o

    def mul_by_cofactor(self, p, out=None):
        """ExtendedPoint::mul_by_cofactor (src/lib.rs:722-724)."""
        return self._call("jj_mul_by_cofactor", [(p, EXT_W, np.uint64)], EXT_W, out=out)

    def point_neg(self, p, out=None):
        """Neg for ExtendedPoint (src/lib.rs:195-210): (-U, V, Z, -T1, T2)."""
        return self._call("jj_point_neg", [(p, EXT_W, np.uint64)], EXT_W, out=out)

    def point_eq(self, p, q):
        """ConstantTimeEq for ExtendedPoint (src/lib.rs:153-163), element by element: flags[i] = (p[i] == q[i])."""
        o = self._call("jj_point_eq", [(p, EXT_W, np.uint64), (q, EXT_W, np.uint64)], 1, np.uint8)
        return o if isinstance(o, DeviceArray) else o.reshape(-1)

    def affine_to_extended(self, a, out=None):
        """From<AffinePoint> for ExtendedPoint (src/lib.rs:214-226): (u, v) -> (u, v, 1, u, v)."""
        return self._call("jj_affine_to_extended", [(a, AFF_W, np.uint64)], EXT_W, out=out)

    def affine_to_bytes(self, a, out=None):
        return self._call("jj_affine_to_bytes", [(a, AFF_W, np.uint64)], 32, np.uint8, out=out)

    def batch_from_bytes(self, enc, zip216=True):
        """AffinePoint::batch_from_bytes (src/lib.rs:541-627) -> (points, ok)."""
        return self._call("jj_batch_from_bytes", [(enc, 32, np.uint8)], AFF_W, ok=True,
                          flags=0 if zip216 else L.JJ_PRE_ZIP216)

    def _flag(self, name, p, flags=0):
        o = self._call(name, [(p, EXT_W, np.uint64)], 1, np.uint8, flags=flags)
        return o if isinstance(o, DeviceArray) else o.reshape(-1)

    def is_torsion_free(self, p, ladder=False):
        """ExtendedPoint::is_torsion_free (src/lib.rs:709-711).  Default: the order-8 pairing test (csrc/torsion.cuh);
        ladder=True decides by the reference's own [r]P == O (the cross-check, ~7x the work)."""
        return self._flag("jj_is_torsion_free", p, L.JJ_TORSION_LADDER if ladder else 0)

    def is_prime_order(self, p):
        """ExtendedPoint::is_prime_order (src/lib.rs:717-719)."""
        return self._flag("jj_is_prime_order", p)

    def is_identity(self, p):
        return self._flag("jj_is_identity", p)

    def is_small_order(self, p):
        return self._flag("jj_is_small_order", p)

    # ---- CUDA graphs ---------------------------------------------------------------------------
    def graph_capture(self, fn):
        """Capture the device-resident, JJ_ASYNC calls made by fn() into a graph; returns a handle for graph_launch.
        Run fn() once eagerly first so that every scratch buffer exists."""
        self._check(self.lib.jj_graph_begin(self.ctx))
        try:
            fn()
        finally:
            g = C.c_void_p()
            rc = self.lib.jj_graph_end(self.ctx, C.byref(g))
        self._check(rc)
        return g

    def graph_launch(self, g):
        self._check(self.lib.jj_graph_launch(self.ctx, g))

    def graph_destroy(self, g):
        self._check(self.lib.jj_graph_destroy(self.ctx, g))

    # ---- measurement helpers -----------------------------------------------------------------
    def sync(self):
        self._check(self.lib.jj_sync(self.ctx))

    def timer_start(self):
        self._check(self.lib.jj_timer_start(self.ctx))

    def timer_stop(self):
        ms = C.c_float()
        self._check(self.lib.jj_timer_stop(self.ctx, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        self._check(self.lib.jj_flush_l2(self.ctx))

    def launch_count(self):
        return int(self.lib.jj_launch_count(self.ctx))

    def imad_peak(self):
        v = C.c_double()
        self._check(self.lib.jj_measure_imad_peak(self.ctx, C.byref(v)))
        return v.value

    def set_scalar_mul_variant(self, v):
        self._check(self.lib.jj_set_scalar_mul_variant(self.ctx, v))

    def device_info(self):
        sm, khz, mem = C.c_int32(), C.c_int32(), C.c_uint64()
        self._check(self.lib.jj_device_info(self.ctx, C.byref(sm), C.byref(khz), C.byref(mem)))
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "hbm_bytes": mem.value}

    # ---- multi-GPU ------------------------------------------------------------------------------
    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        rc = self.lib.jj_comm_unique_id(buf)
        if rc != 0:
            raise JubjubError(rc, "ncclGetUniqueId failed (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def comm_init(self, nranks, rank, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._check(self.lib.jj_comm_init(self.ctx, nranks, rank, buf))
        self.nranks, self.rank = nranks, rank

    def ipc_export(self, darr):
        """64-byte CUDA IPC handle of a DeviceArray (to be sent to the peer processes)."""
        buf = (C.c_char * 64)()
        self._check(self.lib.jj_ipc_export(self.ctx, darr.ptr, buf))
        return bytes(buf)

    def ipc_open(self, handle):
        p = C.c_void_p()
        self._check(self.lib.jj_ipc_open(self.ctx, (C.c_char * 64).from_buffer_copy(handle), C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        self._check(self.lib.jj_ipc_close(self.ctx, ptr))

    def point_sum_sharded(self, points_local, output="extended"):
        """Sum over a batch spread across the ranks (local sums + all-gather of the partial sums): one point, the same on
        every rank."""
        w, dt, f = self._out_fmt(output)
        o = self._out(True, 1, w, dt)
        self._check(self.lib.jj_point_sum_sharded(self.ctx, points_local.ptr, o.ptr, points_local.shape[0],
                                                  f | L.JJ_DEVICE_PTRS))
        return o

    def set_peer_outputs(self, ptrs):
        """Register every rank's gathered-output buffer (rank order; own buffer at index rank) to fuse
        the all-gather into the scalar-mul kernel (P2P stores over NVLink).  None / [] switches back."""
        if not ptrs:
            self._check(self.lib.jj_comm_set_peer_outputs(self.ctx, None, 0))
            return
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self._check(self.lib.jj_comm_set_peer_outputs(self.ctx, arr, len(ptrs)))

    def scalar_mul_sharded_vartime(self, points_local, scalars_local, out_all, output="extended", async_=False,
                                   n_total=None, out_local_host=None):
        """This rank's block + all-gather of every rank's results into out_all (device).  points/scalars hold this
        rank's block only: DeviceArrays, or host arrays (staged in chunks that overlap the kernels).  n_total: size of
        the whole batch when the blocks are ragged (block partition of shard_range); default: equal blocks.
        out_local_host: optional host array receiving this rank's own block of results."""
        w, dt, f = self._out_fmt(output)
        dev = isinstance(points_local, DeviceArray)
        if isinstance(scalars_local, DeviceArray) != dev:
            raise ValueError("points and scalars must both be DeviceArrays or both host arrays")
        f |= (L.JJ_DEVICE_PTRS if dev else 0) | (L.JJ_ASYNC if async_ and dev else 0)
        n_local = points_local.shape[0]
        if scalars_local.shape[0] != n_local:
            raise JubjubError(-1, f"length mismatch: {scalars_local.shape[0]} != {n_local}")
        if out_local_host is not None and (out_local_host.shape != (n_local, w) or out_local_host.dtype != np.dtype(dt)
                                           or not out_local_host.flags["C_CONTIGUOUS"]):
            raise ValueError(f"out_local_host must be a contiguous ({n_local}, {w}) {np.dtype(dt)} array")
        if not dev:
            points_local = np.ascontiguousarray(points_local, dtype=np.uint64)
            scalars_local = np.ascontiguousarray(scalars_local, dtype=np.uint8)
        hp = out_local_host.ctypes.data if out_local_host is not None else None
        if n_total is None:
            if out_local_host is None and dev:
                self._check(self.lib.jj_scalar_mul_sharded(self.ctx, points_local.ptr, scalars_local.ptr, out_all.ptr,
                                                           n_local, f))
                return out_all
            n_total = n_local * self.nranks
        self._check(self.lib.jj_scalar_mul_sharded_n(self.ctx, self._ptr(points_local), self._ptr(scalars_local),
                                                     out_all.ptr, hp, n_total, f))
        return out_all


_default = None


def default_engine():
    global _default
    if _default is None:
        _default = Engine(0)
    return _default
