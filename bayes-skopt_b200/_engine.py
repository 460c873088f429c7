"""Thin host layer over libbgp: one handle + one CUDA stream per estimator.

PyTorch is used for device memory, the stream and events only; every numeric step on the
path is a libbgp kernel.  Nothing here falls back to the CPU."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import Op, Prior, check

_F64 = torch.float64


def nvtx_range(name):
    """NVTX range around a host-side phase ("bgp.<phase>"): shows up on the timeline of any CUDA profiler."""
    return torch.cuda.nvtx.range(name)


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def find_zeroable_white(kernel):
    """The WhiteKernel noise_set_to_zero() switches off: first White that is a direct child of
    a (nested) Sum, k1 before k2 -- skopt's ``_param_for_white_kernel_in_Sum`` as used by
    bask/bayesgpr.py:328-333.  Returns (object, "k1__k2"-style parameter name) or (None, None)."""
    def walk(k, prefix):
        if type(k).__name__ == "Sum":
            for name in ("k1", "k2"):
                child = getattr(k, name)
                if type(child).__name__ == "WhiteKernel":
                    return child, prefix + name
                found = walk(child, prefix + name + "__")
                if found[0] is not None:
                    return found
        return None, None
    return walk(kernel, "")


def compile_kernel(kernel):
    """Walks a scikit-learn style kernel tree (duck-typed: class names and attributes of
    sklearn.gaussian_process.kernels, which skopt's kernels subclass) into libbgp's postfix
    program.  theta slots follow scikit-learn's ordering (k1.theta ++ k2.theta, fixed
    hyper-parameters take no slot; sklearn:kernels.py:738-766)."""
    ops, fixed_ls = [], []
    state = {"theta": 0}
    white_obj, _ = find_zeroable_white(kernel)

    def fixed(bounds):
        return isinstance(bounds, str) and bounds == "fixed"

    def walk(k):
        name = type(k).__name__
        if name in ("Sum", "Product"):
            walk(k.k1)
            walk(k.k2)
            ops.append(Op(_lib.OP_ADD if name == "Sum" else _lib.OP_MUL, -1, 0, 0, 0.0, 0, 0))
        elif name == "Exponentiation":
            walk(k.kernel)
            ops.append(Op(_lib.OP_POW, -1, 0, 0, float(k.exponent), 0, 0))
        elif name == "ConstantKernel":
            if fixed(k.constant_value_bounds):
                ops.append(Op(_lib.OP_CONST, -1, 0, 0, float(k.constant_value), 0, 0))
            else:
                ops.append(Op(_lib.OP_CONST, state["theta"], 0, 0, float(k.constant_value), 0, 0))
                state["theta"] += 1
        elif name == "WhiteKernel":
            flags = _lib.FLAG_ZEROABLE_WHITE if k is white_obj else 0
            if fixed(k.noise_level_bounds):
                ops.append(Op(_lib.OP_WHITE, -1, 0, flags, float(k.noise_level), 0, 0))
            else:
                ops.append(Op(_lib.OP_WHITE, state["theta"], 0, flags, float(k.noise_level), 0, 0))
                state["theta"] += 1
        elif name in ("RBF", "Matern"):
            if name == "RBF":
                code = _lib.OP_RBF
            else:
                nu = float(k.nu)
                code = {0.5: _lib.OP_MATERN12, 1.5: _lib.OP_MATERN32, 2.5: _lib.OP_MATERN52,
                        float("inf"): _lib.OP_RBF}.get(nu)
                if code is None:
                    raise NotImplementedError(
                        f"Matern(nu={nu}) needs the modified Bessel function; libbgp implements "
                        "nu in {0.5, 1.5, 2.5, inf}")
            ls = np.atleast_1d(np.asarray(k.length_scale, dtype=np.float64))
            n_ls = len(ls)
            if fixed(k.length_scale_bounds):
                off = len(fixed_ls)
                if n_ls > 1:
                    fixed_ls.extend(ls.tolist())
                ops.append(Op(code, -1, n_ls, 0, float(ls[0]), off, 0))
            else:
                ops.append(Op(code, state["theta"], n_ls, 0, float(ls[0]), 0, 0))
                state["theta"] += n_ls
        else:
            raise NotImplementedError(f"kernel {name} is not supported by the B200 path")

    walk(kernel)
    if len(ops) > _lib.BGP_MAX_OPS:
        raise NotImplementedError("kernel tree too large for the device program")
    return ops, fixed_ls, state["theta"]


class Factor:
    """Device-resident factorisation(s): the tiled L / L^-1 slabs, z = L^-1 y, LML, info."""

    def __init__(self, thetas, slabs, z, lml, info):
        self.thetas, self.slabs, self.z, self.lml, self.info = thetas, slabs, z, lml, info

    def __len__(self):
        return self.thetas.shape[0]


class Engine:
    def __init__(self, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.BgpError("no CUDA device visible: the B200 path has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.stream = torch.cuda.Stream(device=self.device)
        h = C.c_void_p()
        check(self.lib.bgp_create(C.byref(h), self.device.index), "bgp_create")
        self.h = h
        self.n = self.d = self.p = 0
        self.launches = 0   # kernels enqueued by this engine (bench.py's gpu_launches)
        self._keep = []

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.close_peers()
                self.lib.bgp_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    @property
    def _st(self):
        return C.c_void_p(self.stream.cuda_stream)

    _NP = {torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32, torch.int64: np.int64}

    def _staging(self, nbytes):
        """Next slot of a small ring of reusable pinned host buffers (cudaHostAlloc per copy costs
        ~0.5 ms; the ring makes every H2D a memcpy into pinned memory + one async DMA)."""
        if not hasattr(self, "_ring"):
            self._ring, self._ring_at = [None] * 8, 0
        i = self._ring_at
        self._ring_at = (i + 1) % len(self._ring)
        slot = self._ring[i]
        if slot is not None:
            slot[1].synchronize()          # the previous copy out of this slot has completed
        if slot is None or slot[0].numel() < nbytes:
            cap = max(1 << 16, 1 << int(nbytes - 1).bit_length())
            slot = [torch.empty(cap, dtype=torch.uint8).pin_memory(), torch.cuda.Event()]
            self._ring[i] = slot
        return slot

    def to_dev(self, a, dtype=_F64):
        a = np.ascontiguousarray(a, dtype=self._NP[dtype])
        nbytes = a.nbytes
        if nbytes == 0:
            return self.empty(*a.shape, dtype=dtype)
        buf, ev = self._staging(nbytes)
        buf[:nbytes].numpy()[:] = a.reshape(-1).view(np.uint8)
        with torch.cuda.stream(self.stream):
            dev = buf[:nbytes].to(self.device, non_blocking=True)
            ev.record(self.stream)
        return dev.view(dtype).reshape(a.shape)

    def empty(self, *shape, dtype=_F64):
        with torch.cuda.stream(self.stream):
            return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=_F64):
        with torch.cuda.stream(self.stream):
            return torch.zeros(*shape, dtype=dtype, device=self.device)

    def to_host(self, t):
        self.stream.synchronize()
        return t.cpu().numpy()

    def sync(self):
        self.stream.synchronize()

    def fetch_after(self, event, *tensors):
        """Host copies of ``tensors`` that wait for ``event`` only: the D2H runs on a side stream,
        so work enqueued on the engine's stream after the event keeps the GPU busy meanwhile."""
        if not hasattr(self, "_side"):
            self._side, self._side_pinned = torch.cuda.Stream(device=self.device), {}
        self._side.wait_event(event)
        outs = []
        with torch.cuda.stream(self._side):
            for i, t in enumerate(tensors):
                key = (i, tuple(t.shape), t.dtype)
                h = self._side_pinned.get(key)
                if h is None:
                    h = self._side_pinned[key] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                h.copy_(t, non_blocking=True)
                outs.append(h)
        self._side.synchronize()
        return [h.numpy().copy() for h in outs]

    # ------------------------------------------------------------------ model set-up
    def set_kernel(self, kernel, n_warp=0):
        """Uploads the kernel program.  ``n_warp`` = d switches input warping on: theta rows then
        are ``[kernel theta, log a_1..a_d, log b_1..b_d]`` (bask/bayesgpr.py:351-365) and
        ``self.p`` is the full row length, ``self.p_kernel`` the kernel's own part."""
        ops, fixed_ls, p = compile_kernel(kernel)
        arr = (Op * len(ops))(*ops)
        fl = (C.c_double * max(1, len(fixed_ls)))(*fixed_ls)
        check(self.lib.bgp_set_kernel(self.h, arr, len(ops), p, fl, len(fixed_ls)), "bgp_set_kernel")
        if n_warp:
            check(self.lib.bgp_set_warp(self.h, int(n_warp)), "bgp_set_warp")
        self.p_kernel, self.n_warp = p, int(n_warp)
        self.p = p + 2 * int(n_warp)

    def set_priors(self, table):
        if table is None:
            check(self.lib.bgp_set_priors(self.h, None, 0), "bgp_set_priors")
            return
        arr = (Prior * len(table))()
        for i, (kind, params) in enumerate(table):
            arr[i].kind = kind
            for j, v in enumerate(params):
                arr[i].p[j] = v
        check(self.lib.bgp_set_priors(self.h, arr, len(table)), "bgp_set_priors")

    def set_data(self, X, y, alpha):
        X = np.ascontiguousarray(X, dtype=np.float64)
        n, d = X.shape
        alpha = np.ascontiguousarray(np.broadcast_to(np.asarray(alpha, dtype=np.float64), (n,)))
        # one staging copy and one H2D for the three arrays; bgp_set_data copies them into the handle's
        # own buffers in stream order, so nothing here has to wait for the device
        packed = self.to_dev(np.concatenate([X.reshape(-1), np.asarray(y, dtype=np.float64).reshape(n), alpha]))
        Xd, yd, ad = packed[: n * d], packed[n * d: n * d + n], packed[n * d + n:]
        check(self.lib.bgp_set_data(self.h, _ptr(Xd), _ptr(yd), _ptr(ad), n, d, self._st), "bgp_set_data")
        self.n, self.d = n, d

    # ------------------------------------------------------------------ K1 + K2
    def logprob_dev(self, thetas_dev, lp_extra_dev=None, want_lml=False):
        B = thetas_dev.shape[0]
        lp = self.empty(B)
        lml = self.empty(B) if want_lml else None
        info = self.empty(B, dtype=torch.int32)
        check(self.lib.bgp_logprob_batched(self.h, _ptr(thetas_dev), B, _ptr(lp_extra_dev), _ptr(lp),
                                           _ptr(lml), _ptr(info), self._st), "bgp_logprob_batched")
        self.launches += self._lp_launches()   # (scale_x +) gram + chol, or the fused small-n kernel, per wave
        return lp, lml, info

    def logprob(self, thetas, lp_extra=None):
        th = self.to_dev(np.atleast_2d(thetas))
        ex = None if lp_extra is None else self.to_dev(lp_extra)
        lp, lml, info = self.logprob_dev(th, ex, want_lml=True)
        self.stream.synchronize()
        return lp.cpu().numpy(), lml.cpu().numpy(), info.cpu().numpy()

    def factorize(self, thetas):
        th = thetas if torch.is_tensor(thetas) else self.to_dev(np.atleast_2d(thetas))
        S = th.shape[0]
        slab = int(self.lib.bgp_factor_slab_doubles(self.h))
        slabs = self.empty(S, slab)
        z = self.empty(S, self.n)
        lml = self.empty(S)
        info = self.empty(S, dtype=torch.int32)
        check(self.lib.bgp_factorize_batched(self.h, _ptr(th), S, _ptr(slabs), _ptr(z), _ptr(lml),
                                             _ptr(info), self._st), "bgp_factorize_batched")
        self.launches += 3 if self._lp_launches() == 3 else 2   # (scale_x +) gram + chol
        return Factor(th, slabs, z, lml, info)

    def extract(self, factor, index, what):
        n = self.n
        out = self.empty(n) if what == _lib.EXTRACT_ALPHA else self.empty(n, n)
        check(self.lib.bgp_factor_extract(self.h, _ptr(factor.slabs[index]), _ptr(factor.z[index]), what,
                                          _ptr(out), self._st), "bgp_factor_extract")
        self.launches += 2 if what == _lib.EXTRACT_KINV else 1
        return out

    def lml_gradient(self, theta_row):
        """(LML, dLML/dtheta, info) at one theta row: factorise, extract alpha_ and K_inv_, then the
        analytic-gradient kernel (sklearn:_gpr.py:583-651).  The gradient covers the kernel's own
        hyper-parameters (the first ``p_kernel`` entries of the row)."""
        f = self.factorize(np.atleast_2d(theta_row))
        alpha = self.extract(f, 0, _lib.EXTRACT_ALPHA)
        kinv = self.extract(f, 0, _lib.EXTRACT_KINV)
        grad = self.empty(max(self.p_kernel, 1))
        check(self.lib.bgp_lml_gradient(self.h, _ptr(f.thetas), _ptr(alpha), _ptr(kinv), _ptr(grad), self._st),
              "bgp_lml_gradient")
        self.launches += 3
        self.stream.synchronize()
        return float(f.lml.cpu().numpy()[0]), grad.cpu().numpy()[: self.p_kernel], int(f.info.cpu().numpy()[0])

    # ------------------------------------------------------------------ K4
    def predict(self, factor, Xc_dev, thetas_dev=None, noise_off=True, y_mean=0.0, y_std=1.0,
                zextra=None, want_v=False):
        th = factor.thetas if thetas_dev is None else thetas_dev
        S, m = th.shape[0], Xc_dev.shape[0]
        mu, sd = self.empty(S, m), self.empty(S, m)
        R = 0 if zextra is None else zextra.shape[1]
        dots = self.empty(S, R, m) if R else None
        v_ld = 32 * ((self.n + 31) // 32)
        v = self.empty(S, m, v_ld) if want_v else None
        check(self.lib.bgp_predict_batched(self.h, _ptr(th), S, _ptr(factor.slabs), _ptr(factor.z),
                                           _ptr(Xc_dev), m, 1 if noise_off else 0, float(y_mean),
                                           float(y_std), _ptr(mu), _ptr(sd), _ptr(zextra), R, _ptr(dots),
                                           _ptr(v), v_ld, self._st), "bgp_predict_batched")
        self.launches += 1
        return mu, sd, dots, v

    def acq(self, kind, mu, sd, p0=float("nan"), gumbel32=None, want_fit=False):
        S, m = mu.shape
        per_theta = self.empty(S, m)
        out = self.empty(m)
        skipped = self.empty(S, dtype=torch.int32)
        fit = self.empty(S, 5) if want_fit else None
        K = 0 if gumbel32 is None else gumbel32.shape[1]
        check(self.lib.bgp_acq_sweep(self.h, kind, _ptr(mu), _ptr(sd), S, m, float(p0), _ptr(gumbel32), K,
                                     _ptr(per_theta), _ptr(out), _ptr(skipped), _ptr(fit), self._st),
              "bgp_acq_sweep")
        self.launches += {_lib.ACQ_EI: 4, _lib.ACQ_TTEI: 7, _lib.ACQ_MEAN: 4, _lib.ACQ_LCB: 4,
                          _lib.ACQ_MES: 15}[kind]
        return out, per_theta, skipped, fit

    def argmax(self, v):
        idx = self.empty(1, dtype=torch.int64)
        check(self.lib.bgp_argmax(self.h, _ptr(v), v.shape[0], _ptr(idx), self._st), "bgp_argmax")
        self.launches += 1
        return idx

    # ------------------------------------------------------------------ multi-GPU plumbing
    def connect_peers(self, group, max_walkers):
        """Maps the walker-exchange blocks of all ranks of `group` into this process (cudaIpc over NVLink
        P2P): after this the sharded MCMC needs neither NCCL nor the host between propose and accept.
        torch.distributed only carries the 64-byte handles here."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        peers = getattr(self, "_peers", None)
        if peers is not None and peers == (id(group), world, rank) and max_walkers <= self._peer_cap:
            return
        self.close_peers()
        cap = max(int(max_walkers), 256)
        mine = (C.c_ubyte * _lib.IPC_HANDLE_BYTES)()
        check(self.lib.bgp_peer_export(self.h, cap, mine), "bgp_peer_export")
        with torch.cuda.stream(self.stream):
            t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
            allh = torch.empty(world * _lib.IPC_HANDLE_BYTES, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, t, group=group)
        self.stream.synchronize()
        raw = bytes(allh.cpu().numpy().tobytes())
        check(self.lib.bgp_peer_connect(self.h, raw, rank, world), "bgp_peer_connect")
        dist.barrier(group=group)           # every rank has mapped every block before anyone stores into one
        self._peers, self._peer_cap, self._peer_group = (id(group), world, rank), cap, group
        self._mc_buffers_sharded = None

    def close_peers(self):
        if getattr(self, "_peers", None) is not None:
            self.stream.synchronize()
            self.lib.bgp_peer_close(self.h)
            self._peers = None

    def mcmc_sharded(self, pos, n_steps, seed, group, a=2.0, buffers=None):
        """Walker-sharded run over the ranks of `group` (one CUDA graph per rank, peer stores for the
        log-prob exchange).  Same buffers as ``mcmc``; the chain is identical on every rank."""
        W, p = pos.shape
        self.connect_peers(group, W)
        if buffers is None or buffers["chain"].shape != (n_steps, W, p):
            buffers = dict(pos=self.empty(W, p), lp=self.empty(W), chain=self.empty(n_steps, W, p),
                           lpc=self.empty(n_steps, W), acc=self.empty(W, dtype=torch.int32))
        src = pos if torch.is_tensor(pos) else self.to_dev(pos)
        with torch.cuda.stream(self.stream):
            buffers["pos"].copy_(src, non_blocking=True)
        check(self.lib.bgp_mcmc_run_sharded(self.h, _ptr(buffers["pos"]), _ptr(buffers["lp"]), W, n_steps, float(a),
                                            C.c_uint64(int(seed) & (2 ** 64 - 1)), _ptr(buffers["chain"]),
                                            _ptr(buffers["lpc"]), _ptr(buffers["acc"]), self._st),
              "bgp_mcmc_run_sharded")
        L = self._lp_launches()
        self.launches += 1 + L + 2 + 2 * (L + 1) * n_steps   # exchange + initial log-posterior; the rest as in mcmc()
        return buffers

    def peer_timed_out(self):
        flag = C.c_int(0)
        check(self.lib.bgp_peer_status(self.h, C.byref(flag)), "bgp_peer_status")
        return bool(flag.value)

    def _lp_launches(self):
        """Kernels per batched log-posterior wave for the current model (asked from the library)."""
        return int(self.lib.bgp_logprob_launches(self.h))

    def peer_counters(self):
        """(nanoseconds spent inside peer exchanges, number of exchanges) of this rank so far."""
        out = (C.c_uint64 * 2)()
        check(self.lib.bgp_peer_counters(self.h, out), "bgp_peer_counters")
        return int(out[0]), int(out[1])

    # ------------------------------------------------------------------ K3
    def mcmc(self, pos, n_steps, seed, a=2.0, buffers=None):
        """Device-resident run: returns (pos, lp, chain, lp_chain, accepted) tensors."""
        W, p = pos.shape
        if buffers is None or buffers["chain"].shape != (n_steps, W, p):
            buffers = dict(pos=self.empty(W, p), lp=self.empty(W), chain=self.empty(n_steps, W, p),
                           lpc=self.empty(n_steps, W), acc=self.empty(W, dtype=torch.int32))
        src = pos if torch.is_tensor(pos) else self.to_dev(pos)     # pinned staging ring, engine stream
        with torch.cuda.stream(self.stream):
            buffers["pos"].copy_(src, non_blocking=True)
        check(self.lib.bgp_mcmc_run(self.h, _ptr(buffers["pos"]), _ptr(buffers["lp"]), W, n_steps, float(a),
                                    C.c_uint64(int(seed) & (2 ** 64 - 1)), _ptr(buffers["chain"]),
                                    _ptr(buffers["lpc"]), _ptr(buffers["acc"]), self._st), "bgp_mcmc_run")
        L = self._lp_launches()
        # initial log-posterior, colours of all steps, first proposals; per half step: L + (accept + next proposals)
        self.launches += L + 2 + 2 * (L + 1) * n_steps
        return buffers
