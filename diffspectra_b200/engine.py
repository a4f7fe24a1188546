"""Host-side owner of one ds_ctx: packed weights, molecule plans, workspaces and the calls into the C-ABI.

PyTorch is used here only for device memory, streams and dtype plumbing; all arithmetic of the hot path runs in
libdiffspectra_b200.so.  There is no fallback: without a B200 + the built library every call raises.
"""
import ctypes

import numpy as np
import torch

from . import _lib as L


class Plan:
    """Packed ragged layout of one batch (atoms / unordered pairs)."""

    def __init__(self, engine, n_atoms, N):
        n_atoms = np.ascontiguousarray(np.asarray(n_atoms, dtype=np.int32))
        self.B = int(n_atoms.shape[0])
        self.N = int(N)
        self.n_atoms = n_atoms
        nbytes = L.lib().ds_plan_bytes(self.B, self.N)
        if nbytes == 0:
            raise L.DiffSpectraError('invalid plan shape B=%d N=%d' % (self.B, self.N))
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=engine.device)
        mn, mp = ctypes.c_int(0), ctypes.c_int(0)
        L.check(L.lib().ds_plan_build(engine.h, n_atoms.ctypes.data_as(ctypes.c_void_p), self.B, self.N, L.ptr(self.buf),
                                      ctypes.byref(mn), ctypes.byref(mp), L.stream_ptr()), 'ds_plan_build')
        self.Mn, self.Mp = mn.value, mp.value

    def args(self):
        return L.ptr(self.buf), self.B, self.N, self.Mn, self.Mp


class Engine:
    def __init__(self, device, mode='bf16', spectra_version='allspectra', model_kind='DMT'):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise L.DiffSpectraError('diffspectra_b200 runs on a CUDA (sm_100a) device only; got %s — there is no CPU '
                                     'fallback for the sampling hot path' % self.device)
        self.mode = {'fp32': L.MODE_FP32, 'bf16': L.MODE_BF16}[mode]
        self.mode_name = mode
        self.spectra_version = spectra_version
        self.model_kind = model_kind
        self.h = ctypes.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(idx):
            L.check(L.lib().ds_create_model(ctypes.byref(self.h), idx, self.mode, L.SPECTRA_VERSIONS[spectra_version],
                                            L.MODEL_KINDS[model_kind]), 'ds_create_model')
        self.blob = None
        self._ws = None
        self._spec_ws = None

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                L.lib().ds_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def pack_weights(self, state_dict):
        keep = []
        names, ptrs = [], []
        for k, v in state_dict.items():
            if not torch.is_floating_point(v):
                continue
            t = v.detach().to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(t)
            names.append(k.encode())
            ptrs.append(t.data_ptr())
        n = len(names)
        nbytes = L.lib().ds_packed_weights_bytes(self.h)
        if self.blob is None or self.blob.numel() < nbytes:
            self.blob = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        c_names = (ctypes.c_char_p * n)(*names)
        c_ptrs = (ctypes.c_void_p * n)(*ptrs)
        L.check(L.lib().ds_pack_weights(self.h, c_names, c_ptrs, n, L.ptr(self.blob), ctypes.c_size_t(self.blob.numel()),
                                        L.stream_ptr()), 'ds_pack_weights')
        torch.cuda.current_stream().synchronize()    # `keep` (fp32 staging copies) may be freed after this

    # ------------------------------------------------------------------ buffers
    def plan(self, n_atoms, N=None):
        n_atoms = np.asarray(n_atoms, dtype=np.int32)
        return Plan(self, n_atoms, int(n_atoms.max()) if N is None else N)

    def workspace(self, plan):
        need = L.lib().ds_workspace_bytes(self.h, plan.B, plan.Mn, plan.Mp)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def launch_count(self):
        return int(L.lib().ds_launch_count(self.h))

    # ------------------------------------------------------------------ calls
    def context_embedding(self, context):
        """cond_lin(SpecFormer(context)) [B,1024]; context = Tensor [B,1,L] / [B,L] or list of three."""
        if isinstance(context, (list, tuple)):
            spectra = [self._dev(c).reshape(c.shape[0], -1) for c in context]
            if self.spectra_version != 'allspectra' or len(spectra) != 3:
                raise ValueError('spectra_version should be uv, ir, raman or allspectra')
            uv, ir, raman = spectra
        else:
            sp = self._dev(context).reshape(context.shape[0], -1)
            uv = sp if self.spectra_version == 'uv' else None
            ir = sp if self.spectra_version == 'ir' else None
            raman = sp if self.spectra_version == 'raman' else None
            if uv is None and ir is None and raman is None:
                raise ValueError('spectra_version should be uv, ir, raman or allspectra')
        first = uv if uv is not None else (ir if ir is not None else raman)
        B = first.shape[0]
        lens = {'uv': 701, 'ir': 3501, 'raman': 3501}
        for name, t in (('uv', uv), ('ir', ir), ('raman', raman)):
            if t is not None and tuple(t.shape) != (B, lens[name]):
                raise ValueError('%s spectrum must have shape [B,%d], got %s' % (name, lens[name], tuple(t.shape)))
        need = L.lib().ds_specformer_workspace_bytes(self.h, B)
        if self._spec_ws is None or self._spec_ws.numel() < need:
            self._spec_ws = None
            self._spec_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty(B, 1024, dtype=torch.float32, device=self.device)
        L.check(L.lib().ds_specformer_ctx(self.h, L.ptr(uv), L.ptr(ir), L.ptr(raman), B, L.ptr(out), L.ptr(self._spec_ws),
                                          ctypes.c_size_t(self._spec_ws.numel()), L.stream_ptr()), 'ds_specformer_ctx')
        return out

    def _dev(self, t):
        return t.detach().to(device=self.device, dtype=torch.float32).contiguous()

    def denoise(self, plan, x, edge_x, noise_level, ctx_emb, cond_x=None, cond_edge_x=None):
        B, N = plan.B, plan.N
        x, edge_x, noise_level, ctx_emb = self._dev(x), self._dev(edge_x), self._dev(noise_level), self._dev(ctx_emb)
        assert tuple(x.shape) == (B, N, 9) and tuple(edge_x.shape) == (B, N, N, 2), (x.shape, edge_x.shape)
        if cond_x is not None:
            cond_x, cond_edge_x = self._dev(cond_x), self._dev(cond_edge_x)
        out_x = torch.empty_like(x)
        out_e = torch.empty_like(edge_x)
        ws = self.workspace(plan)
        L.check(L.lib().ds_denoise(self.h, *plan.args(), L.ptr(x), L.ptr(edge_x), L.ptr(cond_x), L.ptr(cond_edge_x),
                                   L.ptr(noise_level), L.ptr(ctx_emb), L.ptr(out_x), L.ptr(out_e), L.ptr(ws),
                                   ctypes.c_size_t(ws.numel()), L.stream_ptr()), 'ds_denoise')
        return out_x, out_e

    def sample_loop(self, plan, ctx_emb, coef_table, z=None, edge_z=None, raw_noise=None, seed=0, gid_base=0,
                    temperature=1.0, use_graph=True, first_step=0, steps=None, out=None):
        """Runs table rows [first_step, first_step+steps).  raw_noise = (raw_pos[S,B,N,3], raw_h[S,B,N,6],
        raw_e[S,B,2,N,N]) for this segment or None (device Philox)."""
        B, N = plan.B, plan.N
        total = coef_table.shape[0]
        steps = total - first_step if steps is None else steps
        ctx_emb, coef_table = self._dev(ctx_emb), self._dev(coef_table)
        if z is not None:
            z, edge_z = self._dev(z), self._dev(edge_z)
        rp = rh = re = None
        if raw_noise is not None:
            rp, rh, re = [self._dev(t) for t in raw_noise]
            assert rp.shape[0] >= steps
        if out is None:
            out = (torch.empty(B, N, 9, dtype=torch.float32, device=self.device),
                   torch.empty(B, N, N, 2, dtype=torch.float32, device=self.device))
        ws = self.workspace(plan)
        L.check(L.lib().ds_sample_loop(self.h, *plan.args(), L.ptr(z), L.ptr(edge_z), L.ptr(ctx_emb), L.ptr(coef_table),
                                       int(first_step), int(steps), L.ptr(rp), L.ptr(rh), L.ptr(re),
                                       ctypes.c_ulonglong(seed), ctypes.c_longlong(gid_base), ctypes.c_float(temperature),
                                       int(bool(use_graph)), L.ptr(out[0]), L.ptr(out[1]), L.ptr(ws),
                                       ctypes.c_size_t(ws.numel()), L.stream_ptr()), 'ds_sample_loop')
        self._keep = (ctx_emb, coef_table, z, edge_z, rp, rh, re)     # alive until the stream has consumed them
        return out

    def sampler_step(self, plan, x, edge_x, pred, edge_pred, coef_row, raw_noise=None, seed=0, gid_base=0, step_index=0,
                     temperature=1.0):
        x, edge_x = self._dev(x).clone(), self._dev(edge_x).clone()
        pred, edge_pred, coef_row = self._dev(pred), self._dev(edge_pred), self._dev(coef_row)
        rp = rh = re = None
        if raw_noise is not None:
            rp, rh, re = [self._dev(t) for t in raw_noise]
        xm, em = torch.empty_like(x), torch.empty_like(edge_x)
        ws = self.workspace(plan)
        L.check(L.lib().ds_sampler_step(self.h, *plan.args(), L.ptr(x), L.ptr(edge_x), L.ptr(pred), L.ptr(edge_pred),
                                        L.ptr(coef_row), L.ptr(rp), L.ptr(rh), L.ptr(re), ctypes.c_ulonglong(seed),
                                        ctypes.c_longlong(gid_base), int(step_index), ctypes.c_float(temperature),
                                        L.ptr(xm), L.ptr(em), L.ptr(ws), ctypes.c_size_t(ws.numel()), L.stream_ptr()),
                'ds_sampler_step')
        return x, edge_x, xm, em

    def post_process(self, plan, x_mean, edge_mean):
        B, N = plan.B, plan.N
        x_mean, edge_mean = self._dev(x_mean), self._dev(edge_mean)
        pos = torch.empty(B, N, 3, dtype=torch.float32, device=self.device)
        atom = torch.empty(B, N, dtype=torch.int32, device=self.device)
        fc = torch.empty(B, N, dtype=torch.int32, device=self.device)
        bond = torch.empty(B, N, N, dtype=torch.float32, device=self.device)
        ws = self.workspace(plan)
        L.check(L.lib().ds_post_process(self.h, *plan.args(), L.ptr(x_mean), L.ptr(edge_mean), L.ptr(pos), L.ptr(atom),
                                        L.ptr(fc), L.ptr(bond), L.ptr(ws), ctypes.c_size_t(ws.numel()), L.stream_ptr()),
                'ds_post_process')
        return pos, atom, fc, bond

    def molecule_records(self, plan, x_mean, edge_mean, rec_n=None):
        """[B, ds_record_bytes(rec_n)] uint8 on the device: pos f32[R*3] | atom u8[R] | fc i8[R] | bond u8[R*R] | n u8, written
        by one kernel from the means of the last step (post_process + mol_process, sampling.py:12-32,53-97)."""
        B, N = plan.B, plan.N
        rec_n = N if rec_n is None else max(int(rec_n), N)
        x_mean, edge_mean = self._dev(x_mean), self._dev(edge_mean)
        rb = int(L.lib().ds_record_bytes(rec_n))
        rec = torch.empty(B, rb, dtype=torch.uint8, device=self.device)
        L.check(L.lib().ds_molecule_records(self.h, *plan.args(), L.ptr(x_mean), L.ptr(edge_mean), rec_n, L.ptr(rec),
                                            ctypes.c_size_t(rec.numel()), L.stream_ptr()), 'ds_molecule_records')
        return rec
