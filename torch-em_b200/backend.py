"""The one compute backend of the product: calls into ``libb200em.so`` (hand-written sm_100a CUDA) through the C ABI.

Each method takes torch tensors only to read ``data_ptr()``, shapes and the channel pitch; the arithmetic is in the
library.  There is no CPU or PyTorch fallback: tensors that are not on a CUDA device raise.
"""
import collections
import ctypes

import torch

from . import _lib
from ._lib import ACT, BF16, F32, call


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"activations must be float32 or bfloat16, got {t.dtype}")


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _act(t):
    """(pointer, pitch) of an NDHWC activation view (N, D, H, W, C) whose base buffer is contiguous."""
    if t.device.type != "cuda":
        raise RuntimeError("b200em: tensors must live on a CUDA device (there is no CPU fallback for this path)")
    assert t.dim() == 5 and t.stride(4) == 1, "NDHWC activation view expected"
    ld = t.stride(3)
    N, D, H, W, C = t.shape
    assert ld >= C and t.stride(2) == W * ld and t.stride(1) == H * W * ld and (N == 1 or t.stride(0) == D * H * W * ld), \
        "activation view must be a channel slice of a contiguous NDHWC buffer"
    return _ptr(t), ld


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _f32(t):
    if t is None:
        return None
    assert t.dtype == torch.float32 and t.is_contiguous()
    return _ptr(t)


class WeightPack:
    """Kernel-operand copies of one conv weight (derived data; the fp32 nn.Parameter stays the master)."""
    __slots__ = ("w_fwd_f32", "w_dgrad_f32", "f32_stale", "umma_fwd", "umma_dgrad", "cout", "cin", "kernel", "thin", "thin_kp",
                 "ds_fwd", "ds_dgrad", "master", "tf32_fwd", "tf32_dgrad", "h16_fwd", "h16_dgrad", "h16_ds_fwd")


class H16Operand:
    """fp16 operand copy of an fp32 activation for the h16 path: ``t`` (N, D, H, W, C) float16, dense, = fp16(2^k * x_hat);
    ``absmax`` the 1-element device tensor k is derived from on the device (None: k = 0, a normalised tensor)."""
    __slots__ = ("t", "absmax")

    def __init__(self, t, absmax):
        self.t, self.absmax = t, absmax


class _PackJob(ctypes.Structure):
    """struct b200em_pack_job of include/b200em.h (64 bytes)."""
    _fields_ = [("w", ctypes.c_void_p), ("packed", ctypes.c_void_p),
                ("Cout", ctypes.c_int32), ("Cin", ctypes.c_int32), ("kd", ctypes.c_int32), ("kh", ctypes.c_int32),
                ("kw", ctypes.c_int32), ("dgrad", ctypes.c_int32), ("layout", ctypes.c_int32),
                ("CC", ctypes.c_int32), ("NPb", ctypes.c_int32), ("block_begin", ctypes.c_int32),
                ("reserved", ctypes.c_int32 * 2)]


PACK_PLAIN, PACK_DEPTH_STACKED, PACK_PLAIN_TF32, PACK_PLAIN_F16, PACK_DEPTH_STACKED_F16 = 0, 1, 2, 3, 4


class PackSet:
    """The operand images of ALL conv weights of one model on one device, rebuilt by ONE launch per direction.

    ``refresh_fwd()`` runs at the start of every forward pass, ``refresh_dgrad()`` at the start of every backward pass
    (csrc/pack.cu).  Rebuilding unconditionally instead of caching by ``tensor._version`` means that in-place updates the
    version counter does not see (``p.data.mul_()``, EMA, old-style optimizers, ``load_state_dict``) can never leave a stale
    operand behind.  The job table lives on the device and is rebuilt only when a parameter's storage moves.
    The set is owned by the model (``model._pack_store``), so it dies with it."""

    def __init__(self, backend, weights):
        self.B = backend
        self.sig = None
        self.packs = {}
        self._build(weights)

    @staticmethod
    def signature(weights):
        return tuple((k, w.data_ptr(), tuple(w.shape), str(w.device)) for k, w in weights.items())

    def _build(self, weights):
        lib = _lib.load()
        B = self.B
        self.sig = self.signature(weights)
        self.packs = {}
        self.device = None
        # table -> [(pack, attr, layout, elems)]: 0 / 1 = bf16 forward / data-gradient operands, 2 / 3 = the fp32 operands of the
        # TF32 path, 4 / 5 = the fp16 operands of the h16 path (each allocated and packed only when a pass really runs with fp32
        # activations on that path)
        plan = {0: [], 1: [], 2: [], 3: [], 4: [], 5: []}
        for key, w in weights.items():
            if w.device.type != "cuda":
                raise RuntimeError("b200em: parameters must live on a CUDA device (there is no CPU fallback for this path)")
            self.device = w.device
            cout, cin, kd, kh, kw = w.shape
            wd = w.detach()
            assert wd.dtype == torch.float32 and wd.is_contiguous()
            pk = WeightPack()
            pk.cout, pk.cin, pk.kernel, pk.master = cout, cin, (kd, kh, kw), wd
            pk.w_fwd_f32 = pk.w_dgrad_f32 = None
            pk.f32_stale = True
            pk.umma_fwd = pk.umma_dgrad = pk.thin = pk.ds_fwd = pk.ds_dgrad = pk.tf32_fwd = pk.tf32_dgrad = None
            pk.h16_fwd = pk.h16_dgrad = pk.h16_ds_fwd = None
            pk.thin_kp = 0
            n = cout * cin * kd * kh * kw
            if B.use_umma:
                # one packed operand per direction: the depth-stacked layout where that kernel takes the layer, else the plain one
                if B.use_ds and lib.b200em_conv3d_umma_ds_supported(cin, cout, kd, kh, kw):
                    plan[0].append((pk, "ds_fwd", PACK_DEPTH_STACKED, n))
                elif lib.b200em_conv3d_umma_supported(cin, cout, kd, kh, kw):
                    plan[0].append((pk, "umma_fwd", PACK_PLAIN, n))
                if B.use_ds and lib.b200em_conv3d_umma_ds_supported(cout, cin, kd, kh, kw):
                    plan[1].append((pk, "ds_dgrad", PACK_DEPTH_STACKED, n))
                elif lib.b200em_conv3d_umma_supported(cout, cin, kd, kh, kw):
                    plan[1].append((pk, "umma_dgrad", PACK_PLAIN, n))
                if lib.b200em_conv3d_umma_tf32_supported(cin, cout, kd, kh, kw):
                    plan[2].append((pk, "tf32_fwd", PACK_PLAIN_TF32, n))
                if lib.b200em_conv3d_umma_tf32_supported(cout, cin, kd, kh, kw):
                    plan[3].append((pk, "tf32_dgrad", PACK_PLAIN_TF32, n))
                if lib.b200em_conv3d_umma_supported(cin, cout, kd, kh, kw):
                    plan[4].append((pk, "h16_fwd", PACK_PLAIN_F16, n))
                if B.use_ds and lib.b200em_conv3d_umma_ds_supported(cin, cout, kd, kh, kw):
                    plan[4].append((pk, "h16_ds_fwd", PACK_DEPTH_STACKED_F16, n))     # forward only: the depth-stacked kernel, fp32 out
                if lib.b200em_conv3d_umma_supported(cout, cin, kd, kh, kw):
                    plan[5].append((pk, "h16_dgrad", PACK_PLAIN_F16, n))
                kp = -(-kd * kh * kw * cin // 32) * 32
                first = B.use_ds and lib.b200em_conv3d_first_supported(cin, cout, kd, kh, kw)
                if cin <= 4 and not first and lib.b200em_conv3d_umma_supported(kp, cout, 1, 1, 1):
                    pk.thin_kp = kp                   # thin-K first conv (im2col layout), packed by refresh_fwd
            self.packs[key] = pk
        self.plan = plan
        self.tables = {}
        for d in (0, 1):
            self._build_table(d)

    def _build_table(self, d):
        plan = self.plan[d]
        if not plan:
            self.tables[d] = None
            return
        dtype = (torch.bfloat16, torch.float32, torch.float16)[d // 2]
        total = sum(-(-n // 8) * 8 for _, _, _, n in plan)
        flat = torch.empty(total, dtype=dtype, device=self.device)
        jobs = (_PackJob * len(plan))()
        off = 0
        for i, (pk, attr, layout, n) in enumerate(plan):
            buf = flat[off:off + n]
            off += -(-n // 8) * 8                     # keep every image 16-byte aligned
            setattr(pk, attr, buf)
            j = jobs[i]
            j.w, j.packed = pk.master.data_ptr(), buf.data_ptr()
            j.Cout, j.Cin = pk.cout, pk.cin
            j.kd, j.kh, j.kw = pk.kernel
            j.dgrad, j.layout = d % 2, layout
        nblocks = ctypes.c_int(0)
        call("b200em_pack_batch_prepare", ctypes.cast(jobs, ctypes.c_void_p), len(plan), ctypes.byref(nblocks))
        host = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8)
        self.tables[d] = (host.to(self.device), len(plan), nblocks.value, flat)

    def _launch(self, d):
        if d not in self.tables:
            self._build_table(d)
        t = self.tables[d]
        if t is None:
            return
        table, njobs, nblocks, _ = t
        with torch.cuda.device(self.device):
            call("b200em_pack_batch", _ptr(table), njobs, nblocks, ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))

    def refresh_fwd(self, weights=None, bf16=True):
        """Rebuild the forward operands from the CURRENT parameter values (one launch).  ``weights``: the model's current
        {key: weight}; if a parameter's storage moved (``.to()``, re-assignment) the job table is rebuilt first."""
        if weights is not None and self.signature(weights) != self.sig:
            self._build(weights)
        for pk in self.packs.values():
            pk.f32_stale = True
        if not bf16:
            # fp32 activations: the h16 (default) or TF32 tensor-core path when torch allows TF32 convolutions, else the exact
            # direct kernels (their fp32 operands are packed lazily, f32_operands)
            if self.B.h16_enabled():
                self._launch(4)
            elif self.B.tf32_enabled():
                self._launch(2)
            return
        self._launch(0)
        for pk in self.packs.values():
            if pk.thin_kp:
                self._pack_thin(pk)

    def refresh_dgrad(self, bf16=True):
        if bf16:
            self._launch(1)
        elif self.B.h16_enabled():
            self._launch(5)
        elif self.B.tf32_enabled():
            self._launch(3)

    def _pack_thin(self, pk):
        # thin-K first conv: W'[co][tap*Cin+ci] = W[co][ci][tap], zero padded to Kp channels (im2col layout)
        wd, cout, cin = pk.master, pk.cout, pk.cin
        taps = pk.kernel[0] * pk.kernel[1] * pk.kernel[2]
        wt = torch.zeros((cout, pk.thin_kp), dtype=torch.float32, device=wd.device)
        wt[:, :taps * cin] = wd.reshape(cout, cin, taps).permute(0, 2, 1).reshape(cout, taps * cin)
        if pk.thin is None:
            pk.thin = torch.empty(cout * pk.thin_kp, dtype=torch.bfloat16, device=wd.device)
        with torch.cuda.device(wd.device):
            call("b200em_conv3d_umma_pack", _ptr(wt), cout, pk.thin_kp, 1, 1, 1, 0, _ptr(pk.thin), _stream(wd))

    def __getitem__(self, key):
        return self.packs[key]


# One scratch buffer per device for the whole process: the library keeps the raw pointer (b200em_set_workspace), so the tensor must
# never be freed or replaced while any backend object may still launch kernels on that device.
_WORKSPACES = {}


class CudaBackend:
    name = "cuda"

    def __init__(self, use_umma=True, use_ds=True, use_cs=True, use_tf32=None, use_h16=True, use_splitk=True):
        self.use_umma = use_umma
        self.use_tf32 = use_tf32    # None: follow torch.backends.cudnn.allow_tf32 (True by default, like the reference's fp32 runs)
        self.use_splitk = use_splitk
        self.use_h16 = use_h16      # TF32-class arithmetic through fp16 operand copies at the bf16 MMA rate (False: kind::tf32 kernels)
        self._force_h16 = False     # fp16 autocast in effect (forcing_h16)
        self._h16_recent = []       # [(tensor, H16Operand)]: the last two un-normalised fp32 tensors converted (dz feeds wgrad AND dgrad)
        self._absmax_known = []     # [(fp32 tensor, its device max |.|)] written by the producing kernel, consumed by to_h16
        self._workspaces = _WORKSPACES   # device index -> zero-filled scratch registered with the library (split-K partial sums)
        self.use_ds = use_ds and use_umma
        self.use_cs = use_cs and use_umma
        self.timing = None          # {family: [(start_event, end_event, work), ...]} while bench.py measures
        self.calls = collections.Counter()   # conv launches per "<kernel>:<direction>" (tests assert which kernels ran)

    # ---- per-kernel timing (bench.py roofline leg) -----------------------------------------------------------
    def start_timing(self):
        self.timing = {}

    def stop_timing(self):
        """-> {family: (launches, total ms, total work)}; call after torch.cuda.synchronize()."""
        out = {}
        for fam, evs in (self.timing or {}).items():
            out[fam] = (len(evs), sum(a.elapsed_time(b) for a, b, _ in evs), sum(w for _, _, w in evs))
        self.timing = None
        return out

    def _timed(self, family, work, fn):
        self.calls[family] += 1
        if self.timing is None:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        self.timing.setdefault(family, []).append((a, b, work))
        return r

    def tf32_enabled(self):
        """fp32 activations take the TF32 tensor-core kernels iff torch would run an fp32 convolution in TF32
        (``torch.backends.cudnn.allow_tf32``, default True); set it to False for the exact-fp32 CUDA-core kernels."""
        if not self.use_umma:
            return False
        return torch.backends.cudnn.allow_tf32 if self.use_tf32 is None else bool(self.use_tf32)

    WORKSPACE_BYTES = 65 << 20

    def ensure_workspace(self, device):
        """Registers (once per device) the zero-filled scratch the conv kernels use to split the reduction of layers with fewer
        output tiles than SMs (b200em_set_workspace).  The kernels leave it all-zero; it is used by one stream at a time -- the
        schedule runs every conv of a device on that device's current stream."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._workspaces:
            with torch.cuda.device(idx):
                ws = torch.zeros(self.WORKSPACE_BYTES, dtype=torch.uint8, device=torch.device("cuda", idx))
                call("b200em_set_workspace", _ptr(ws), ws.numel())
            self._workspaces[idx] = ws

    def h16_enabled(self):
        """fp32 activations with TF32 allowed run as fp16 operand copies (same 11-bit significand as TF32, round-to-nearest, exact
        power-of-two range scaling) on kind::f16 MMAs with fp32 accumulation: TF32-class results at twice the kind::tf32 rate."""
        if self._force_h16 and self.use_umma:
            return True                      # fp16 autocast: half-precision operands are what the caller asked for
        return bool(self.use_h16) and self.tf32_enabled()

    def forcing_h16(self, on=True):
        """Context manager: run fp32 activations on the h16 path whatever ``allow_tf32`` says -- how the model serves
        ``torch.autocast(float16)`` (the reference trainer's default mixed precision): fp32 activations and accumulation, fp16
        tensor-core operands, i.e. at least the precision fp16 autocast would give."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            was = self._force_h16
            self._force_h16 = bool(on) or was
            try:
                yield
            finally:
                self._force_h16 = was
        return ctx()

    def _want_absmax(self, out):
        """Device word for max |out| of a gradient tensor the h16 path will convert next (None when it will not)."""
        if out is None or out.dtype != torch.float32 or not self.h16_enabled() or out.shape[4] % 8:
            return None
        am = torch.zeros(1, dtype=torch.float32, device=out.device)
        self._absmax_known = [e for e in self._absmax_known if e[0] is not out][-3:] + [(out, am)]   # a rewrite supersedes
        return am

    def to_h16(self, x, in_ss, colsum=None):
        """fp16 operand copy of the fp32 NDHWC view ``x``: fp16(scale * x + shift) for a normalised tensor (unit scale: no range
        problem), else fp16(2^k * x) with k taken on the device from max |x| (gradients span any range).  Un-normalised tensors are
        remembered by identity (the same dz object feeds the weight gradient and the data gradient back to back).
        colsum (C floats, fp32): += the per-channel sums of x from the same pass (the bias gradient, from the fp32 values)."""
        if in_ss is None:
            for t, h in self._h16_recent:
                if t is x:
                    if colsum is not None:
                        s2 = torch.zeros((x.shape[0], x.shape[4], 2), dtype=torch.float32, device=x.device)
                        self.channel_sums(x, s2)
                        colsum += s2[:, :, 0].sum(0)
                    return h
        N, D, H, W, C = x.shape
        xp, xld = _act(x)
        out = torch.empty((N, D, H, W, C), dtype=torch.float16, device=x.device)
        am = None
        if in_ss is None:
            # max |x| already came out of the kernel that produced x (norm / max-pool / head backward), else one more pass
            for i, (t, a) in enumerate(self._absmax_known):
                if t is x:
                    am = a
                    del self._absmax_known[i]
                    break
            if am is None:
                am = torch.zeros(1, dtype=torch.float32, device=x.device)
                call("b200em_absmax_f32", xp, xld, N * D * H * W, C, _f32(am), None, _stream(x))
        call("b200em_cvt_f16", xp, xld, _f32(in_ss), _f32(am), _ptr(out), _f32(colsum) if C <= 8192 else None, N, D * H * W, C, _stream(x))
        if colsum is not None and C > 8192:
            s2 = torch.zeros((N, C, 2), dtype=torch.float32, device=x.device)
            self.channel_sums(x, s2)
            colsum += s2[:, :, 0].sum(0)
        h = H16Operand(out, am)
        if in_ss is None:
            self._h16_recent = self._h16_recent[-1:] + [(x, h)]
        return h

    # ---- weights ---------------------------------------------------------------------------------------------
    def pack_set(self, weights):
        """{key: conv weight} -> PackSet (all operand images of a model, one launch per direction)."""
        return PackSet(self, weights)

    def pack(self, key, w, lazy_dgrad=False):
        """Operand images of ONE conv weight, both directions, packed now (tests / stand-alone use of conv())."""
        ps = PackSet(self, {key: w})
        ps.refresh_fwd()
        ps.refresh_dgrad()
        if self.tf32_enabled():
            ps._launch(2)
            ps._launch(3)
        if self.h16_enabled():
            ps._launch(4)
            ps._launch(5)
        return ps[key]

    def f32_operands(self, pk):
        """fp32 (taps, Cin, Cout) / (taps, Cout, Cin) operands of the direct CUDA-core kernels, packed on first use after
        every refresh of the pack set."""
        if pk.f32_stale:
            pk.f32_stale = False
            cout, cin = pk.cout, pk.cin
            kd, kh, kw = pk.kernel
            taps = kd * kh * kw
            wd = pk.master
            if pk.w_fwd_f32 is None:
                pk.w_fwd_f32 = torch.empty((taps, cin, cout), dtype=torch.float32, device=wd.device)
                pk.w_dgrad_f32 = torch.empty((taps, cout, cin), dtype=torch.float32, device=wd.device)
            with torch.cuda.device(wd.device):
                call("b200em_pack_conv_weights", _ptr(wd), cout, cin, kd, kh, kw, _ptr(pk.w_fwd_f32), _ptr(pk.w_dgrad_f32),
                     _stream(wd))
        return pk.w_fwd_f32, pk.w_dgrad_f32

    # ---- layout / statistics ---------------------------------------------------------------------------------
    def to_ndhwc(self, x, y):
        N, C, D, H, W = x.shape
        yp, yld = _act(y)
        call("b200em_ncdhw_to_ndhwc", _f32(x), yp, _dt(y), yld, N, C, D * H * W, _stream(x))

    def channel_sums(self, x, sums):
        N, D, H, W, C = x.shape
        xp, xld = _act(x)
        call("b200em_channel_sums", xp, xld, _dt(x), N, D * H * W, C, _f32(sums), _stream(x))

    def channel_dot_sums(self, g, x, sums):
        N, D, H, W, C = x.shape
        gp, gld = _act(g)
        xp, xld = _act(x)
        call("b200em_channel_dot_sums", gp, gld, xp, xld, _dt(x), N, D * H * W, C, _f32(sums), _stream(x))

    def norm_finalize(self, sums, S, groups, gamma, beta, eps):
        N, C, _ = sums.shape
        ss = torch.empty((N, C, 2), dtype=torch.float32, device=sums.device)
        mr = torch.empty((N, C, 2), dtype=torch.float32, device=sums.device)
        g = gamma.detach() if gamma is not None else None
        b = beta.detach() if beta is not None else None
        call("b200em_norm_finalize", _f32(sums), N, C, S, groups, _f32(g), _f32(b), eps, _f32(ss), _f32(mr), _stream(sums))
        return ss, mr

    def norm_bwd_finalize(self, dsums, mr, gamma, S, groups, dgamma, dbeta):
        N, C, _ = dsums.shape
        coef = torch.empty((N, C, 3), dtype=torch.float32, device=dsums.device)
        g = gamma.detach() if gamma is not None else None
        call("b200em_norm_bwd_finalize", _f32(dsums), _f32(mr), _f32(g), N, C, S, groups, _f32(coef), _f32(dgamma),
             _f32(dbeta), _stream(dsums))
        return coef

    def norm_bwd_apply(self, g, x, coef, add, out, relu_mask):
        N, D, H, W, C = g.shape
        gp, gld = _act(g)
        xp, xld = _act(x)
        ap, ald = _act(add) if add is not None else (None, 0)
        op, old = _act(out)
        call("b200em_norm_bwd_apply", gp, gld, xp, xld, _f32(coef), ap, ald, op, old, _dt(g), N, D * H * W, C,
             int(relu_mask), _f32(self._want_absmax(out)), _stream(g))

    # ---- convolution -----------------------------------------------------------------------------------------
    def conv(self, x, in_ss, pack, bias, y, sums, kernel, relu, dgrad, dot_x=None):
        """y = act(conv(norm(x)) + bias).  dgrad=True: the data-gradient (flipped/transposed weights).
        sums += (sum y, sum y^2) per (n, channel), or (sum y, sum y * dot_x) when dot_x is given (norm backward)."""
        N, D, H, W, Cin = x.shape
        Cout = y.shape[4]
        xp, xld = _act(x)
        yp, yld = _act(y)
        assert _dt(x) == _dt(y)
        if self.use_umma and self.use_splitk:
            self.ensure_workspace(x.device)
        b = bias.detach() if bias is not None else None
        kd, kh, kw = kernel
        flops = 2.0 * N * D * H * W * Cin * Cout * kd * kh * kw
        if (not dgrad) and self.use_ds and Cin == 1 and xld == 1 and x.dtype == torch.bfloat16 and yld % 8 == 0 and \
                y.data_ptr() % 16 == 0 and dot_x is None and _lib.load().b200em_conv3d_first_supported(Cin, Cout, kd, kh, kw):
            # first conv of the network: im2col rows built on the fly in shared memory, straight from the fp32 parameter
            self._timed("first:fwd", flops, lambda: call(
                "b200em_conv3d_first", xp, _f32(in_ss), _f32(pack.master), _f32(b), yp, yld, _f32(sums), N, D, H, W, Cout,
                int(relu), _stream(x)))
            return None
        if (not dgrad) and self.use_ds and Cin == 1 and xld == 1 and x.dtype == torch.float32 and self.h16_enabled() and yld % 4 == 0 and \
                y.data_ptr() % 16 == 0 and dot_x is None and _lib.load().b200em_conv3d_first_supported(Cin, Cout, kd, kh, kw):
            # first conv with fp32 activations on the h16 path: the same kernel with an fp16 im2col image and fp32 output
            self._timed("first:fwd", flops, lambda: call(
                "b200em_conv3d_first_f32", xp, _f32(in_ss), _f32(pack.master), _f32(b), yp, yld, _f32(sums), N, D, H, W, Cout,
                int(relu), _stream(x)))
            return None
        if (not dgrad) and pack.thin is not None and x.dtype == torch.bfloat16 and yld % 8 == 0 and y.data_ptr() % 16 == 0:
            cols = self.im2col(x, in_ss, kernel, pack.thin_kp)
            assert dot_x is None
            self._timed("thin:fwd", flops, lambda: call(
                "b200em_conv3d_umma", _ptr(cols), pack.thin_kp, None, _ptr(pack.thin), _f32(b), yp, yld, _f32(sums), None, 0, N, D, H,
                W, pack.thin_kp, Cout, 1, 1, 1, int(relu), _stream(x)))
            return cols       # kept by the schedule for the weight gradient of the same conv
        ok16 = x.dtype == torch.bfloat16 and xld % 8 == 0 and yld % 8 == 0 and x.data_ptr() % 16 == 0 and y.data_ptr() % 16 == 0
        wds = pack.ds_dgrad if dgrad else pack.ds_fwd
        if wds is not None and ok16:
            dp, dld = _act(dot_x) if dot_x is not None else (None, 0)
            if dot_x is None or (dld % 8 == 0 and dot_x.data_ptr() % 16 == 0):
                self._timed("ds:dgrad" if dgrad else "ds:fwd", flops, lambda: call(
                    "b200em_conv3d_umma_ds", xp, xld, _f32(in_ss), _ptr(wds), _f32(b), yp, yld, _f32(sums), dp, dld, N, D, H, W,
                    Cin, Cout, kd, kh, kw, int(relu), _stream(x)))
                return None
        wu = pack.umma_dgrad if dgrad else pack.umma_fwd
        if wu is not None and x.dtype == torch.bfloat16 and xld % 8 == 0 and yld % 8 == 0 and \
                x.data_ptr() % 16 == 0 and y.data_ptr() % 16 == 0:
            dp, dld = _act(dot_x) if dot_x is not None else (None, 0)
            if dot_x is None or (dld % 8 == 0 and dot_x.data_ptr() % 16 == 0):
                self._timed("plain:dgrad" if dgrad else "plain:fwd", flops, lambda: call(
                    "b200em_conv3d_umma", xp, xld, _f32(in_ss), _ptr(wu), _f32(b), yp, yld, _f32(sums), dp, dld, N, D, H, W, Cin,
                    Cout, kd, kh, kw, int(relu), _stream(x)))
                return None
        wh = pack.h16_dgrad if dgrad else pack.h16_fwd
        if wh is not None and x.dtype == torch.float32 and self.h16_enabled() and xld % 4 == 0 and yld % 4 == 0 and \
                x.data_ptr() % 16 == 0 and y.data_ptr() % 16 == 0:
            dp, dld = _act(dot_x) if dot_x is not None else (None, 0)
            if dot_x is None or (dld % 4 == 0 and dot_x.data_ptr() % 16 == 0):
                xh = self.to_h16(x, in_ss)
                if (not dgrad) and pack.h16_ds_fwd is not None and xh.absmax is None and dot_x is None and self.use_ds:
                    # few output channels, normalised (unscaled) input: the depth-stacked kernel on the fp16 copy, fp32 output
                    self._timed("h16ds:fwd", flops, lambda: call(
                        "b200em_conv3d_umma_ds_h16", _ptr(xh.t), Cin, _ptr(pack.h16_ds_fwd), _f32(b), yp, yld, _f32(sums), N, D, H, W,
                        Cin, Cout, kd, kh, kw, int(relu), _stream(x)))
                    return xh
                self._timed("h16:dgrad" if dgrad else "h16:fwd", flops, lambda: call(
                    "b200em_conv3d_umma_h16", _ptr(xh.t), Cin, _f32(xh.absmax), _ptr(wh), _f32(b), yp, yld, _f32(sums), dp, dld, N, D,
                    H, W, Cin, Cout, kd, kh, kw, int(relu), _stream(x)))
                return None if dgrad else xh        # kept by the schedule for the weight gradient of the same conv
        wt = pack.tf32_dgrad if dgrad else pack.tf32_fwd
        if wt is not None and x.dtype == torch.float32 and self.tf32_enabled() and xld % 4 == 0 and yld % 4 == 0 and \
                x.data_ptr() % 16 == 0 and y.data_ptr() % 16 == 0:
            dp, dld = _act(dot_x) if dot_x is not None else (None, 0)
            if dot_x is None or (dld % 4 == 0 and dot_x.data_ptr() % 16 == 0):
                self._timed("tf32:dgrad" if dgrad else "tf32:fwd", flops, lambda: call(
                    "b200em_conv3d_umma_tf32", xp, xld, _f32(in_ss), _ptr(wt), _f32(b), yp, yld, _f32(sums), dp, dld, N, D, H, W, Cin,
                    Cout, kd, kh, kw, int(relu), _stream(x)))
                return None
        w = self.f32_operands(pack)[1 if dgrad else 0]
        self._timed("direct:dgrad" if dgrad else "direct:fwd", flops, lambda: call(
            "b200em_conv3d_direct", xp, xld, _f32(in_ss), _f32(w), _f32(b), yp, yld, None if dot_x is not None else _f32(sums),
            _dt(x), N, D, H, W, Cin, Cout, kd, kh, kw, int(relu), _stream(x)))
        if dot_x is not None:
            self.channel_dot_sums(y, dot_x, sums)
        return None

    def im2col(self, x, in_ss, kernel, kp):
        """(N,D,H,W,Cin<=4) -> (N,D,H,W,kp) bf16 im2col of the first conv's taps, with the norm apply fused."""
        N, D, H, W, Cin = x.shape
        xp, xld = _act(x)
        cols = torch.empty((N, D, H, W, kp), dtype=torch.bfloat16, device=x.device)
        kd, kh, kw = kernel
        call("b200em_im2col_taps", xp, xld, _f32(in_ss), _dt(x), _ptr(cols), N, D, H, W, Cin, kd, kh, kw, kp, _stream(x))
        return cols

    def wgrad(self, x, in_ss, dz, dw, db, kernel, aux=None):
        """dw += sum dz * norm(x) shifted by the taps;  db (nullable) += sum dz  (weight and bias gradient).
        aux: what conv() returned for the same (x, in_ss) in the forward pass (the first conv's im2col), or None."""
        N, D, H, W, Cin = x.shape
        Cout = dz.shape[4]
        xp, xld = _act(x)
        zp, zld = _act(dz)
        kd, kh, kw = kernel
        flops = 2.0 * N * D * H * W * Cin * Cout * kd * kh * kw
        taps = kd * kh * kw
        kp = -(-taps * Cin // 32) * 32
        if self.use_umma and self.use_ds and Cin == 1 and xld == 1 and x.dtype == torch.bfloat16 and zld % 8 == 0 and \
                dz.data_ptr() % 16 == 0 and _lib.load().b200em_conv3d_first_supported(Cin, Cout, kd, kh, kw):
            self._timed("first:wgrad", flops, lambda: call(
                "b200em_conv3d_first_wgrad", xp, _f32(in_ss), zp, zld, _f32(dw), _f32(db), N, D, H, W, Cout, _stream(x)))
            return
        if self.use_umma and Cin <= 4 and x.dtype == torch.bfloat16 and zld % 8 == 0 and dz.data_ptr() % 16 == 0 and \
                _lib.load().b200em_conv3d_wgrad_umma_supported(kp, Cout, 1, 1, 1):
            cols = aux if aux is not None else self.im2col(x, in_ss, kernel, kp)
            dwt = torch.zeros((Cout, kp), dtype=torch.float32, device=x.device)
            self._timed("thin:wgrad", flops, lambda: call(
                "b200em_conv3d_wgrad_umma", _ptr(cols), kp, None, zp, zld, _f32(dwt), _f32(db), N, D, H, W, kp, Cout, 1, 1, 1,
                _stream(x)))
            dw += dwt[:, :taps * Cin].reshape(Cout, taps, Cin).permute(0, 2, 1).reshape(dw.shape)
            return
        ok16 = x.dtype == torch.bfloat16 and xld % 8 == 0 and zld % 8 == 0 and x.data_ptr() % 16 == 0 and dz.data_ptr() % 16 == 0
        if self.use_umma and self.use_cs and ok16 and _lib.load().b200em_conv3d_wgrad_cs_supported(Cin, Cout, kd, kh, kw):
            self._timed("cs:wgrad", flops, lambda: call(
                "b200em_conv3d_wgrad_cs", xp, xld, _f32(in_ss), zp, zld, _f32(dw), _f32(db), N, D, H, W, Cin, Cout, kd, kh,
                kw, _stream(x)))
            return
        if self.use_umma and x.dtype == torch.bfloat16 and xld % 8 == 0 and zld % 8 == 0 and x.data_ptr() % 16 == 0 and \
                dz.data_ptr() % 16 == 0 and _lib.load().b200em_conv3d_wgrad_umma_supported(Cin, Cout, kd, kh, kw):
            self._timed("umma:wgrad", flops, lambda: call(
                "b200em_conv3d_wgrad_umma", xp, xld, _f32(in_ss), zp, zld, _f32(dw), _f32(db), N, D, H, W, Cin, Cout, kd, kh,
                kw, _stream(x)))
            return
        if self.h16_enabled() and x.dtype == torch.float32 and xld % 4 == 0 and zld % 4 == 0 and x.data_ptr() % 16 == 0 and \
                dz.data_ptr() % 16 == 0 and Cin % 8 == 0 and Cout % 8 == 0:
            # fp32 activations, TF32 allowed: ONE pass of the tensor-core weight-gradient kernels on fp16 operand copies
            lib = _lib.load()
            fn = "b200em_conv3d_wgrad_cs_h16" if self.use_cs and lib.b200em_conv3d_wgrad_cs_supported(Cin, Cout, kd, kh, kw) else \
                ("b200em_conv3d_wgrad_umma_h16" if lib.b200em_conv3d_wgrad_umma_supported(Cin, Cout, kd, kh, kw) else None)
            if fn is not None:
                xh = aux if isinstance(aux, H16Operand) else self.to_h16(x, in_ss)
                zh = self.to_h16(dz, None, colsum=db)       # bias gradient from the fp32 dz, in the absmax pass
                self._timed("h16:wgrad", flops, lambda: call(
                    fn, _ptr(xh.t), Cin, _f32(xh.absmax), _ptr(zh.t), Cout, _f32(zh.absmax), _f32(dw), None, N, D, H, W, Cin, Cout,
                    kd, kh, kw, _stream(x)))
                return
        first_f32 = Cin == 1 and xld == 1 and self.use_ds and _lib.load().b200em_conv3d_first_supported(Cin, Cout, kd, kh, kw)
        if self.tf32_enabled() and x.dtype == torch.float32 and zld % 4 == 0 and dz.data_ptr() % 16 == 0 and Cout % 8 == 0 and \
                (first_f32 or (xld % 4 == 0 and x.data_ptr() % 16 == 0 and Cin % 8 == 0 and
                               _lib.load().b200em_conv3d_wgrad_umma_supported(Cin, Cout, kd, kh, kw))):
            # fp32 activations, TF32 allowed: the bf16 tensor-core weight-gradient kernels on split operands -- x_hat = hi + lo,
            # dz = hi + lo, dW = hi*hi + hi*lo + lo*hi (the dropped lo*lo term is 2^-16 relative; more accurate than TF32's 10
            # mantissa bits, at three bf16 passes)
            S = D * H * W
            xh = torch.empty((2, N, D, H, W, Cin), dtype=torch.bfloat16, device=x.device)
            zh = torch.empty((2, N, D, H, W, Cout), dtype=torch.bfloat16, device=x.device)
            if first_f32:                                 # the 1-channel network input: a tiny tensor, split with torch ops
                xa = x.float() if in_ss is None else x * in_ss[:, 0, 0].reshape(N, 1, 1, 1, 1) + in_ss[:, 0, 1].reshape(N, 1, 1, 1, 1)
                xh[0].copy_(xa)
                xh[1].copy_(xa - xh[0].float())
            else:
                call("b200em_split_bf16", xp, xld, _f32(in_ss), _ptr(xh[0]), _ptr(xh[1]), N, S, Cin, _stream(x))
            call("b200em_split_bf16", zp, zld, None, _ptr(zh[0]), _ptr(zh[1]), N, S, Cout, _stream(x))

            def three_passes():
                calls, timing, self.timing = self.calls.copy(), self.timing, None
                try:
                    self.wgrad(xh[0], None, zh[0], dw, db, kernel)
                    self.wgrad(xh[0], None, zh[1], dw, db, kernel)
                    self.wgrad(xh[1], None, zh[0], dw, None, kernel)
                finally:
                    self.calls, self.timing = calls, timing    # the three bf16 passes count as ONE fp32 weight gradient

            self._timed("split3:wgrad", flops, three_passes)
            return
        if Cin <= 4:
            self._timed("smallcin:wgrad", flops, lambda: call(
                "b200em_conv3d_wgrad_smallcin", xp, xld, _f32(in_ss), zp, zld, _dt(x), _f32(dw), _f32(db), N, D, H, W, Cin, Cout,
                kd, kh, kw, _stream(x)))
            return
        self._timed("direct:wgrad", flops, lambda: call(
            "b200em_conv3d_wgrad_direct", xp, xld, _f32(in_ss), zp, zld, _dt(x), _f32(dw), N, D, H, W, Cin, Cout, kd, kh, kw,
            _stream(x)))
        if db is not None:
            s = torch.zeros((N, Cout, 2), dtype=torch.float32, device=dz.device)
            self.channel_sums(dz, s)
            db += s[:, :, 0].sum(0)

    # ---- pool / upsample -------------------------------------------------------------------------------------
    def maxpool_fwd(self, x, y, f, sums):
        N, D, H, W, C = x.shape
        assert tuple(y.shape[1:4]) == (D // f[0], H // f[1], W // f[2]), "maxpool_fwd: y must be x's shape floor-divided by the factors"
        xp, xld = _act(x)
        yp, yld = _act(y)
        call("b200em_maxpool3d_fwd", xp, xld, yp, yld, _dt(x), N, D, H, W, C, f[0], f[1], f[2], _f32(sums), _stream(x))

    @staticmethod
    def _coef(coef, C):
        """(pointer, sample stride) of an (N, C, 3) fp32 coefficient view that may be a channel slice of a wider tensor."""
        if coef is None:
            return None, 0
        assert coef.dtype == torch.float32 and coef.shape[1] == C and coef.shape[2] == 3 and coef.stride(2) == 1 and coef.stride(1) == 3
        return _ptr(coef), coef.stride(0)

    def maxpool_bwd(self, x, dp, add, out, f, relu_mask, coef=None):
        """coef (N, C, 3): the added term is c0 * add + c1 * x + c2 (the consuming block's norm backward, fused)."""
        N, D, H, W, C = x.shape
        xp, xld = _act(x)
        dpp, dpld = _act(dp)
        ap, ald = _act(add) if add is not None else (None, 0)
        op, old = _act(out)
        cp, cns = self._coef(coef, C)
        call("b200em_maxpool3d_bwd", xp, xld, dpp, dpld, ap, ald, cp, cns, op, old, _dt(x), N, D, H, W, C, f[0], f[1], f[2],
             int(relu_mask), _f32(self._want_absmax(out)), _stream(x))

    def upsample_fwd(self, x, y, f, sums):
        N, D, H, W, C = x.shape
        assert tuple(y.shape[1:4]) == (D * f[0], H * f[1], W * f[2]), "upsample_fwd: y must be x's shape times the factors"
        xp, xld = _act(x)
        yp, yld = _act(y)
        call("b200em_upsample_trilinear_fwd", xp, xld, yp, yld, _dt(x), N, D, H, W, C, f[0], f[1], f[2], _f32(sums),
             _stream(x))

    def fused_up_bwd_ok(self, dy, f):
        """Whether upsample_bwd can apply the norm backward on the fly (factors (1|2, 2, 2), 16-byte channel vectors whose count
        divides 256, samples below 4 GiB)."""
        vec = 8 if dy.dtype == torch.bfloat16 else 4
        C = dy.shape[4]
        per_sample_bytes = dy.shape[1] * dy.shape[2] * dy.shape[3] * dy.stride(3) * dy.element_size()
        return f[0] <= 2 and f[1] == 2 and f[2] == 2 and C % vec == 0 and 256 % (C // vec) == 0 and dy.stride(3) % vec == 0 and \
            dy.data_ptr() % 16 == 0 and per_sample_bytes < 2 ** 32        # (the kernel addresses a sample with 32-bit byte offsets)

    def upsample_bwd(self, dy, dx, f, zlow=None, coef=None):
        """coef (N, C, 3) + zlow (the tensor that was up-sampled): the gradient that is transposed-interpolated is
        c0 * dy + c1 * up(zlow) + c2 (the consuming block's norm backward, fused; the up-sampled tensor itself is not read)."""
        N, D, H, W, C = dx.shape
        assert tuple(dy.shape[1:4]) == (D * f[0], H * f[1], W * f[2]), "upsample_bwd: dy must be dx's shape times the factors"
        yp, yld = _act(dy)
        xp, xld = _act(dx)
        cp, cns = self._coef(coef, C)
        zp, zld = _act(zlow) if coef is not None else (None, 0)
        call("b200em_upsample_trilinear_bwd", yp, yld, zp, zld, cp, cns, xp, xld, _dt(dx), N, D, H, W, C, f[0], f[1], f[2],
             _stream(dx))

    # ---- head ------------------------------------------------------------------------------------------------
    def head_fwd(self, x, w, b, out, act):
        N, D, H, W, Cin = x.shape
        Cout = out.shape[1]
        xp, xld = _act(x)
        call("b200em_head_fwd", xp, xld, _dt(x), _f32(w.detach().reshape(Cout, Cin)), _f32(b.detach()), _f32(out), N,
             D * H * W, Cin, Cout, ACT[act], _stream(x))

    def head_bwd(self, grad_out, out, x, w, dx, dw, db, act, relu_mask):
        N, D, H, W, Cin = x.shape
        Cout = out.shape[1]
        xp, xld = _act(x)
        dxp, dxld = _act(dx) if dx is not None else (None, 0)
        call("b200em_head_bwd", _f32(grad_out), _f32(out), xp, xld, _dt(x), _f32(w.detach().reshape(Cout, Cin)), dxp, dxld,
             _f32(dw), _f32(db), N, D * H * W, Cin, Cout, ACT[act], int(relu_mask), _f32(self._want_absmax(dx)), _stream(x))


_default = None


def default_backend():
    global _default
    if _default is None:
        _lib.load()
        import os
        _default = CudaBackend(use_ds=os.environ.get("B200EM_DS", "1") == "1", use_cs=os.environ.get("B200EM_CS", "1") == "1",
                               use_h16=os.environ.get("B200EM_H16", "1") == "1", use_splitk=os.environ.get("B200EM_SPLITK", "1") == "1")
    return _default
