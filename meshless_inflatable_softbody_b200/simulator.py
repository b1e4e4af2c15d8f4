"""Simulator: the reference's script-level control functions on one object.

The reference has no class; its public surface is module-level functions and arrays
(sim.py:279-308 setters, :261-266 startup, :341-372 diff_sim, position[f].numpy()).  The
method names and argument meaning below are those; the arithmetic runs in the CUDA library
behind include/mis.h.  There is no CPU path here: without a CUDA device and the built
library, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import native
from .config import SceneConfig


def hash_grid_dims(x0: np.ndarray, h: float):
    """sim.py:123-125: int(2 * (max - min) / float(h) / 3) per axis (float(h) = the fp32 value)."""
    hf = float(np.float32(h))
    p = np.asarray(x0, dtype=np.float64)
    return tuple(max(1, int(2 * (p[:, a].max() - p[:, a].min()) / hf / 3)) for a in range(3))


class Simulator:
    def __init__(self, x0, config: Optional[SceneConfig] = None, device: str = "cuda:0",
                 lanes_per_particle: int = 0, keep_fields: bool = False, graph_steps: int = 0,
                 two_pass_deform: bool = False, cluster_size: int = 0,
                 apply_defaults: bool = True, precision: str = "f32"):
        if not torch.cuda.is_available():
            raise RuntimeError("meshless_inflatable_softbody_b200.Simulator needs a CUDA device (sm_100a); "
                               "there is no CPU fallback")
        self.cfg = config or SceneConfig()
        self.device = torch.device(device)
        self.L = native.lib()
        if precision not in ("f32", "f64"):
            raise ValueError("precision must be 'f32' (sim.py, real = wp.float32) or 'f64' (sim_taichi.py, real = ti.f64)")
        self.f64 = precision == "f64"
        x0_64 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64).reshape(-1, 3)) if self.f64 else None
        x0_np = np.ascontiguousarray(np.asarray(x0, dtype=np.float32).reshape(-1, 3))
        self.n = int(x0_np.shape[0])
        self.x0 = torch.from_numpy(x0_np).to(self.device)
        gx, gy, gz = hash_grid_dims(x0_np, self.cfg.h)
        c = self.cfg
        p = native.MisParams()
        p.h, p.damping, p.dt = c.h, c.damping, c.time_step
        p.k_col, p.col_range = c.collision_penalty_stiffness, c.collision_range
        p.stiff_a, p.stiff_b, p.tanh_k = c.stiffness_a, c.stiffness_b, c.tanh_k
        p.grid_x, p.grid_y, p.grid_z = gx, gy, gz
        p.symmetric_pair, p.identity_rot = int(c.symmetric_pair), int(c.identity_rotation)
        p.self_density, p.euler, p.no_contact = int(c.self_density), int(c.euler), int(not c.ground_contact)
        p.lanes_per_particle, p.keep_fields, p.graph_steps = int(lanes_per_particle), int(keep_fields), int(graph_steps)
        p.two_pass_deform = int(two_pass_deform)
        p.cluster_size = int(cluster_size)
        p.fp64 = int(self.f64)
        self.params = p
        self.hash_grid = (gx, gy, gz)
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream(device=self.device)
        self._h = C.c_void_p()
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        native.check(self.L.mis_create(self.n, self.x0.data_ptr(), C.byref(p), self._st(), C.byref(self._h)), "mis_create")
        self.frame = 0
        if self.f64:            # the neighbour structure was binned on the fp32 rounding of x0; the state and the constants are doubles
            self._set64(0, x0_64, 3)
            k = (C.c_double * 8)(c.h, c.damping, c.time_step, c.collision_penalty_stiffness, c.collision_range, c.stiffness_a, c.stiffness_b, c.tanh_k)
            native.check(self.L.mis_set_constants_f64(self._h, k), "mis_set_constants_f64")
        if apply_defaults:      # main(), sim.py:441-444 + x.fill_(-1.), sim.py:99
            self.set_all_external_force(c.external_force)
            self.set_youngs_modulus(c.youngs_modulus)
            self.set_poisson_ratio(c.poisson_ratio)
            self.set_mass(c.mass)
            self.set_design(c.design_x)

    # ------------------------------------------------------------------ plumbing
    def _st(self):
        return C.c_void_p(self.stream.cuda_stream)

    def _dev(self, a, shape, own: bool = False):
        """Broadcast a scalar / sequence / array / tensor to a contiguous fp32 device tensor the library's stream may read.

        own=True returns a private copy (the setters keep it).  Every kernel that produces the tensor (conversion, expand,
        clone) is enqueued on torch's current stream BEFORE the library stream is made to wait for that stream, and the
        tensor is recorded on the library stream so the caching allocator does not hand its memory out again while a
        library kernel still reads it (torch streams are non-blocking: nothing else orders the two)."""
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=torch.float32)
        else:
            t = torch.as_tensor(np.asarray(a, dtype=np.float32), device=self.device)
        t = t.expand(shape).contiguous()
        if own and isinstance(a, torch.Tensor) and t.data_ptr() == a.data_ptr():
            t = t.clone()
        self._publish(t)
        return t

    def _publish(self, *tensors):
        """Order the library stream after everything enqueued so far on torch's current stream; keep `tensors` alive for it."""
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        for t in tensors:
            if t is not None and t.is_cuda:
                t.record_stream(self.stream)

    # fp64 scenes: doubles in and out (include/mis.h: mis_set_f64 / mis_get_f64)
    _F64 = {"x0": 0, "mass": 1, "youngs": 2, "poisson": 3, "design": 4, "fext": 5, "free": 6, "x": 7, "v": 8, "fel": 9, "vol": 10,
            "rho": 11, "F": 12, "S": 13, "R": 14, "A": 15}

    def _set64(self, what: int, a, dim: int):
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=torch.float64)
        else:
            t = torch.as_tensor(np.asarray(a, dtype=np.float64), device=self.device)
        t = t.expand((self.n, dim) if dim > 1 else (self.n,)).contiguous()
        self._publish(t)
        native.check(self.L.mis_set_f64(self._h, int(what), t.data_ptr(), self._st()), "mis_set_f64")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def _get64(self, what: int, dim) -> torch.Tensor:
        shape = (self.n,) if dim == 1 else ((self.n, 3) if dim == 3 else (self.n, 3, 3))
        out = torch.empty(shape, device=self.device, dtype=torch.float64)
        native.check(self.L.mis_get_f64(self._h, int(what), out.data_ptr(), self._st()), "mis_get_f64")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.synchronize()
            self.L.mis_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self.stream.synchronize()

    # ------------------------------------------------------------------ control functions (sim.py:279-308)
    def set_all_external_force(self, f: Sequence[float]):
        if self.f64:
            return self._set64(5, f, 3)
        self._fext = self._dev(f, (self.n, 3), own=True)
        native.check(self.L.mis_set_ext_force(self._h, self._fext.data_ptr(), self._st()), "mis_set_ext_force")

    def set_external_force(self, i, f):
        """sim.py:279-280; i may be an index or an index array."""
        if self.f64:
            cur = self._get64(5, 3)
            cur[i] = torch.as_tensor(np.asarray(f, np.float64), device=self.device)
            return self._set64(5, cur, 3)
        self._fext[i] = torch.as_tensor(np.asarray(f, np.float32), device=self.device)
        self._publish(self._fext)
        native.check(self.L.mis_set_ext_force(self._h, self._fext.data_ptr(), self._st()), "mis_set_ext_force")

    def set_external_forces(self, f):
        if self.f64:
            return self._set64(5, f, 3)
        self._fext = self._dev(f, (self.n, 3), own=True)
        native.check(self.L.mis_set_ext_force(self._h, self._fext.data_ptr(), self._st()), "mis_set_ext_force")

    def set_external_forces_host(self, f_host: torch.Tensor):
        """Per-step input path: (n,3) fp32 pinned host tensor, copied inside the library."""
        native.check(self.L.mis_set_ext_force_host(self._h, f_host.data_ptr(), self._st()), "mis_set_ext_force_host")

    def set_dirichlet(self, i, d):
        """sim.py:285-286: free_points[i] = d (component-wise multiplier)."""
        if self.f64:
            cur = self._get64(6, 3)
            cur[i] = torch.as_tensor(np.asarray(d, np.float64), device=self.device)
            return self._set64(6, cur, 3)
        if not hasattr(self, "_free"):
            self._free = torch.ones((self.n, 3), device=self.device, dtype=torch.float32)
        self._free[i] = torch.as_tensor(np.asarray(d, np.float32), device=self.device)
        self._publish(self._free)
        native.check(self.L.mis_set_dirichlet(self._h, self._free.data_ptr(), self._st()), "mis_set_dirichlet")

    def set_youngs_modulus(self, E):
        if self.f64:
            self._set64(2, E, 1)
            if not hasattr(self, "_nu64"):
                self._set64(3, 0.0, 1)        # sim_taichi.py:259-264 reads poisson_ratio before it is set: zero-initialised field
            self._E64 = True
            return
        self._E = self._dev(E, (self.n,), own=True)
        if hasattr(self, "_nu"):
            native.check(self.L.mis_set_material(self._h, self._E.data_ptr(), self._nu.data_ptr(), self._st()), "mis_set_material")

    def set_poisson_ratio(self, nu):
        if self.f64:
            self._nu64 = True
            return self._set64(3, nu, 1)
        self._nu = self._dev(nu, (self.n,), own=True)
        if hasattr(self, "_E"):
            native.check(self.L.mis_set_material(self._h, self._E.data_ptr(), self._nu.data_ptr(), self._st()), "mis_set_material")

    def set_mass(self, m):
        if self.f64:
            return self._set64(1, m, 1)
        self._m = self._dev(m, (self.n,), own=True)
        native.check(self.L.mis_set_mass(self._h, self._m.data_ptr(), self._st()), "mis_set_mass")

    def set_design(self, x):
        """x -> ratio = 0.5 tanh(k x) + 0.5 (compute_ratio, sim.py:107-110)."""
        if self.f64:
            return self._set64(4, x, 1)
        self._x = self._dev(x, (self.n,), own=True)
        native.check(self.L.mis_set_design(self._h, self._x.data_ptr(), self._st()), "mis_set_design")

    # ------------------------------------------------------------------ rollout (sim.py:341-358)
    def startup(self, v0: Optional[Sequence[float]] = None):
        self._v0 = tuple(v0) if v0 is not None else tuple(self.cfg.initial_velocity)
        if self.f64:
            native.check(self.L.mis_startup_f64(self._h, (C.c_double * 3)(*self._v0), self._st()), "mis_startup_f64")
            self.frame = 0
            return
        v = (C.c_float * 3)(*self._v0)
        native.check(self.L.mis_startup(self._h, v, self._st()), "mis_startup")
        self.frame = 0

    def set_state(self, x, v, frame: int = 0):
        if self.f64:
            self._set64(7, x, 3); self._set64(8, v, 3)
            self.frame = frame
            return
        xd, vd = self._dev(x, (self.n, 3)), self._dev(v, (self.n, 3))
        native.check(self.L.mis_set_state(self._h, xd.data_ptr(), vd.data_ptr(), self._st()), "mis_set_state")
        self.stream.synchronize()
        self.frame = frame

    def step(self, n_steps: int = 1):
        native.check(self.L.mis_step(self._h, int(n_steps), self._st()), "mis_step")
        self.frame += int(n_steps)

    def rebuild_neighbors(self):
        native.check(self.L.mis_build_neighbors(self._h, self._st()), "mis_build_neighbors")

    # ------------------------------------------------------------------ halo plumbing (slab.py)
    def gather_next_positions(self, ids: torch.Tensor) -> torch.Tensor:
        """New positions x(f+1) (already written by the fused part_1) of the particles `ids` (int32, caller ids)."""
        out = torch.empty((int(ids.numel()), 3), device=self.device, dtype=torch.float32)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        native.check(self.L.mis_gather_next_positions(self._h, ids.data_ptr(), int(ids.numel()), out.data_ptr(), self._st()),
                     "mis_gather_next_positions")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return out

    def scatter_next_positions(self, ids: torch.Tensor, x: torch.Tensor):
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        native.check(self.L.mis_scatter_next_positions(self._h, ids.data_ptr(), int(ids.numel()), x.data_ptr(), self._st()),
                     "mis_scatter_next_positions")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)     # keeps `x` alive until the scatter has read it

    def volumes(self) -> torch.Tensor:
        """V_i = m_i / rho_i (compute_v_i, sim.py:154-167), caller order."""
        return self.fields(("vol",))["vol"]

    def set_volumes(self, ids: torch.Tensor, vol: torch.Tensor):
        """Overwrite V of the particles `ids` (int32 caller ids) -- ghosts of a slab partition take their owner's value."""
        ids = ids.to(device=self.device, dtype=torch.int32).contiguous()
        vol = vol.to(device=self.device, dtype=torch.float32).contiguous()
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        native.check(self.L.mis_set_volumes(self._h, ids.data_ptr(), int(ids.numel()), vol.data_ptr(), self._st()), "mis_set_volumes")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def slots_of(self, ids: torch.Tensor) -> torch.Tensor:
        """Cell-sorted slot of each caller id (the address a peer pushes a ghost's position to)."""
        ids = ids.to(device=self.device, dtype=torch.int32).contiguous()
        out = torch.empty(int(ids.numel()), device=self.device, dtype=torch.int32)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        native.check(self.L.mis_export_slots(self._h, ids.data_ptr(), int(ids.numel()), out.data_ptr(), self._st()), "mis_export_slots")
        self.stream.synchronize()
        return out

    def halo_ipc_handles(self) -> bytes:
        buf = C.create_string_buffer(192)
        native.check(self.L.mis_halo_ipc_handles(self._h, buf), "mis_halo_ipc_handles")
        return bytes(buf.raw)

    def halo_local_ptrs(self):
        a, b, f = C.c_void_p(), C.c_void_p(), C.c_void_p()
        native.check(self.L.mis_halo_local_ptrs(self._h, C.byref(a), C.byref(b), C.byref(f)), "mis_halo_local_ptrs")
        return int(a.value), int(b.value), int(f.value)

    def halo_connect(self, peer_xv0, peer_xv1, peer_flag, push_ids, push_peer, push_slot, ghost_ids, ghost_layer=None):
        """Switch on the fused halo push (include/mis.h, mis_halo_connect).  Pointer lists are ints (device addresses
        valid in this process); push_* / ghost_ids are int32 arrays."""
        k = len(peer_xv0)
        arr = lambda v: (C.c_void_p * max(1, k))(*[C.c_void_p(int(x)) for x in v])
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(np.asarray(a, np.int32)), device=self.device)
        pid, pp, ps, gid = dev(push_ids), dev(push_peer), dev(push_slot), dev(ghost_ids)
        gl = dev(ghost_layer) if ghost_layer is not None else None
        assert gl is None or gl.numel() == gid.numel()
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        native.check(self.L.mis_halo_connect(self._h, k, arr(peer_xv0), arr(peer_xv1), arr(peer_flag),
                                             int(pid.numel()), pid.data_ptr(), pp.data_ptr(), ps.data_ptr(),
                                             int(gid.numel()), gid.data_ptr(), gl.data_ptr() if gl is not None else None,
                                             self._st()), "mis_halo_connect")

    def halo_set_wait(self, wait: bool):
        """wait=False: the per-step flag kernel only publishes; the caller orders the ranks by host synchronisation."""
        native.check(self.L.mis_halo_set_wait(self._h, int(bool(wait))), "mis_halo_set_wait")

    def halo_disconnect(self):
        native.check(self.L.mis_halo_disconnect(self._h), "mis_halo_disconnect")

    def halo_status(self):
        """(timed_out, exchanges): synchronises the stream."""
        e, x = C.c_int(0), C.c_longlong(0)
        native.check(self.L.mis_halo_status(self._h, self._st(), C.byref(e), C.byref(x)), "mis_halo_status")
        return bool(e.value), int(x.value)

    # ------------------------------------------------------------------ obstacle contact (extension)
    def set_sdf_obstacle(self, sdf, bbox_model, xform=None, fd_eps: float = 1e-3):
        """Per-step contact against a DeepSDF-encoded obstacle (extension; the reference evaluates its SDF once,
        sim.py:100, and only the ground plane per step, sim.py:238-244).  Generalises compute_collision_penalty:
        delta = collision_range - sdf(p_model), f = delta^2 * stiffness * n.  bbox_model = (min xyz, max xyz) of the
        model-space region where sdf can be below collision_range; xform = deepsdf.world_to_model_xform(...) or None
        (obstacle given in world coordinates).  sdf=None removes the obstacle."""
        if sdf is None:
            native.check(self.L.mis_set_sdf_contact(self._h, None, None, None, 0.0, self._st()), "mis_set_sdf_contact")
            self._sdf = None
            return
        xf = None if xform is None else (C.c_float * 12)(*[float(v) for v in xform])
        bb = (C.c_float * 6)(*[float(v) for v in np.asarray(bbox_model, np.float32).reshape(-1)])
        self.stream.wait_stream(sdf.stream)
        native.check(self.L.mis_set_sdf_contact(self._h, sdf._h, xf, bb, float(fd_eps), self._st()), "mis_set_sdf_contact")
        self._sdf = sdf          # keep the network alive while the scene uses it

    def contact_force(self) -> torch.Tensor:
        """Obstacle-contact force at the current frame's positions, (n,3), caller order."""
        f = torch.empty((self.n, 3), device=self.device, dtype=torch.float32)
        native.check(self.L.mis_get_contact_force(self._h, f.data_ptr(), self._st()), "mis_get_contact_force")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return f

    def contact_counts(self):
        """(broad-phase candidates, particles in the contact band) of the most recent step."""
        c = (C.c_int * 2)(0, 0)
        native.check(self.L.mis_get_contact_count(self._h, self._st(), c), "mis_get_contact_count")
        return int(c[0]), int(c[1])

    def contact_count(self) -> int:
        return self.contact_counts()[0]

    # ------------------------------------------------------------------ state export
    def position_velocity(self):
        if self.f64:
            return self._get64(7, 3), self._get64(8, 3)
        x = torch.empty((self.n, 3), device=self.device, dtype=torch.float32)
        v = torch.empty((self.n, 3), device=self.device, dtype=torch.float32)
        native.check(self.L.mis_get_state(self._h, x.data_ptr(), v.data_ptr(), self._st()), "mis_get_state")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return x, v

    def position(self) -> torch.Tensor:
        """position[f] in the caller's particle order (sim.py:334)."""
        return self.position_velocity()[0]

    def velocity(self) -> torch.Tensor:
        return self.position_velocity()[1]

    def get_state_host(self, x_host: torch.Tensor, v_host: Optional[torch.Tensor] = None):
        native.check(self.L.mis_get_state_host(self._h, x_host.data_ptr(), v_host.data_ptr() if v_host is not None else None,
                                               self._st()), "mis_get_state_host")

    def get_state_host_async(self, x_host: torch.Tensor, v_host: Optional[torch.Tensor] = None):
        """Streaming export into pinned host tensors: returns at once; the copy overlaps the following steps.
        The tensors are valid after wait_state_host() (at most two exports may be in flight)."""
        native.check(self.L.mis_get_state_host_async(self._h, x_host.data_ptr(), v_host.data_ptr() if v_host is not None else None,
                                                     self._st()), "mis_get_state_host_async")

    def wait_state_host(self, pending_allowed: int = 0):
        """Block until at most `pending_allowed` (0 or 1) streaming exports are still in flight."""
        native.check(self.L.mis_wait_state_host(self._h, int(pending_allowed)), "mis_wait_state_host")

    def fields(self, want=("R", "F", "S", "fel", "rho", "vol")):
        if self.f64:
            dims = {"A": 9, "R": 9, "F": 9, "S": 9, "fel": 3, "rho": 1, "vol": 1}
            return {k: self._get64(self._F64[k], dims[k]) for k in want}
        out = {}
        shapes = {"A": (self.n, 3, 3), "R": (self.n, 3, 3), "F": (self.n, 3, 3), "S": (self.n, 3, 3),
                  "fel": (self.n, 3), "rho": (self.n,), "vol": (self.n,)}
        ptr = {}
        for k in ("A", "R", "F", "S", "fel", "rho", "vol"):
            if k in want:
                out[k] = torch.empty(shapes[k], device=self.device, dtype=torch.float32)
                ptr[k] = out[k].data_ptr()
            else:
                ptr[k] = None
        native.check(self.L.mis_get_fields(self._h, ptr["A"], ptr["R"], ptr["F"], ptr["S"], ptr["fel"], ptr["rho"], ptr["vol"],
                                           self._st()), "mis_get_fields")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return out

    def eval_forces(self, x) -> torch.Tensor:
        xd = self._dev(x, (self.n, 3))
        f = torch.empty((self.n, 3), device=self.device, dtype=torch.float32)
        native.check(self.L.mis_eval_forces(self._h, xd.data_ptr(), f.data_ptr(), self._st()), "mis_eval_forces")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self.stream.synchronize()
        return f

    def export_targets(self, folder: str, every: Optional[int] = None, count: Optional[int] = None):
        """sim.py:363-369: position_{i}.npy / velocity_{i}.npy for i = 1..count at frames every*i.

        Runs the rollout from the current frame; the files are (n,3) fp32 in caller order.
        """
        count = count or self.cfg.target_frames
        every = every or (self.cfg.frames // self.cfg.target_frames)
        os.makedirs(folder, exist_ok=True)
        for i in range(1, count + 1):
            target = every * i
            if target > self.frame:
                self.step(target - self.frame)
            x, v = self.position_velocity()
            np.save(os.path.join(folder, f"position_{i}.npy"), x.cpu().numpy())
            np.save(os.path.join(folder, f"velocity_{i}.npy"), v.cpu().numpy())

    def accumulate_loss(self, target_x, target_v, loss: torch.Tensor):
        """compute_loss (sim.py:269-273) at the current frame: loss[0] (fp64 device tensor) += sum |x - xt|^2 + dt |v - vt|^2."""
        tx, tv = self._dev(target_x, (self.n, 3)), self._dev(target_v, (self.n, 3))
        assert loss.dtype == torch.float64 and loss.is_cuda
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        native.check(self.L.mis_accumulate_loss(self._h, tx.data_ptr(), tv.data_ptr(), loss.data_ptr(), self._st()), "mis_accumulate_loss")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)     # keeps tx / tv alive until the kernel has read them

    def rollout_grad(self, targets, frames: Optional[int] = None, checkpoint_every: int = 0):
        """diff_sim(compute_grad=True) (sim.py:341-372): (loss, d loss / d design x).  targets: a folder of position_{i}.npy /
        velocity_{i}.npy (sim.py:118-119) or a list of (x, v) arrays; target i is compared at frame (frames // len(targets)) * (i + 1).
        The reverse pass (wp.Tape.backward in the reference) is the hand-written adjoint of mis_ref.cuh with checkpointed
        recomputation instead of 3 001 stored frames.  The gradient is float64 for fp64 scenes, float32 otherwise."""
        if isinstance(targets, (str, os.PathLike)):
            k, pairs = 1, []
            while os.path.exists(os.path.join(targets, f"position_{k}.npy")):
                pairs.append((np.load(os.path.join(targets, f"position_{k}.npy")), np.load(os.path.join(targets, f"velocity_{k}.npy"))))
                k += 1
            targets = pairs
        frames = frames or self.cfg.frames
        nt = len(targets)
        tx = torch.as_tensor(np.stack([np.asarray(t[0], np.float32).reshape(self.n, 3) for t in targets]) if nt else np.zeros((1, self.n, 3), np.float32),
                             device=self.device).contiguous()
        tv = torch.as_tensor(np.stack([np.asarray(t[1], np.float32).reshape(self.n, 3) for t in targets]) if nt else np.zeros((1, self.n, 3), np.float32),
                             device=self.device).contiguous()
        self._publish(tx, tv)
        loss = C.c_double(0.0)
        g32 = None if self.f64 else torch.zeros(self.n, device=self.device, dtype=torch.float32)
        g64 = torch.zeros(self.n, device=self.device, dtype=torch.float64) if self.f64 else None
        self.startup(getattr(self, "_v0", None))
        native.check(self.L.mis_rollout_grad(self._h, int(frames), int(nt), tx.data_ptr(), tv.data_ptr(), int(checkpoint_every), C.byref(loss),
                                             g32.data_ptr() if g32 is not None else None, g64.data_ptr() if g64 is not None else None, self._st()),
                     "mis_rollout_grad")
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self.frame = 0
        return float(loss.value), (g64 if self.f64 else g32)

    def rollout_loss(self, targets, frames: Optional[int] = None) -> float:
        """Forward value of the reference's objective (diff_sim without the tape, sim.py:341-362): startup, `frames` steps, and
        compute_loss against target i at frame (frames // len(targets)) * (i + 1).  targets: a folder holding
        position_{i}.npy / velocity_{i}.npy (i = 1..), as export_targets writes them, or a list of (x, v) arrays."""
        if isinstance(targets, (str, os.PathLike)):
            k, pairs = 1, []
            while os.path.exists(os.path.join(targets, f"position_{k}.npy")):
                pairs.append((np.load(os.path.join(targets, f"position_{k}.npy")), np.load(os.path.join(targets, f"velocity_{k}.npy"))))
                k += 1
            targets = pairs
        frames = frames or self.cfg.frames
        every = frames // len(targets)
        loss = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.startup()
        for tx, tv in targets:
            self.step(every)
            self.accumulate_loss(tx, tv, loss)
        self.stream.synchronize()
        return float(loss.item())

    # ------------------------------------------------------------------ neighbour structure (bit-exact checks)
    def neighbor_info(self) -> native.MisNeighborInfo:
        info = native.MisNeighborInfo()
        native.check(self.L.mis_get_neighbor_info(self._h, C.byref(info)), "mis_get_neighbor_info")
        return info

    def cells(self):
        ci = torch.empty(self.n, device=self.device, dtype=torch.int32)
        cc = torch.empty((self.n, 3), device=self.device, dtype=torch.int32)
        pm = torch.empty(self.n, device=self.device, dtype=torch.int32)
        native.check(self.L.mis_export_cells(self._h, ci.data_ptr(), cc.data_ptr(), pm.data_ptr(), self._st()), "mis_export_cells")
        self.stream.synchronize()
        return ci, cc, pm

    def cell_ranges(self):
        info = self.neighbor_info()
        nc = info.cell_dim[0] * info.cell_dim[1] * info.cell_dim[2]
        s = torch.empty(nc, device=self.device, dtype=torch.int32)
        e = torch.empty(nc, device=self.device, dtype=torch.int32)
        native.check(self.L.mis_export_cell_ranges(self._h, s.data_ptr(), e.data_ptr(), self._st()), "mis_export_cell_ranges")
        self.stream.synchronize()
        return s, e

    def neighbors(self):
        info = self.neighbor_info()
        off = torch.empty(self.n + 1, device=self.device, dtype=torch.int64)
        nb = torch.empty(max(1, info.total_pairs), device=self.device, dtype=torch.int32)
        native.check(self.L.mis_export_neighbors(self._h, off.data_ptr(), nb.data_ptr(), self._st()), "mis_export_neighbors")
        self.stream.synchronize()
        return off, nb[: info.total_pairs]

    def profile_step(self, n_steps: int):
        """Device milliseconds summed over n_steps for (deform kernel, force kernel); advances the state."""
        a, b = C.c_double(0), C.c_double(0)
        native.check(self.L.mis_profile_step(self._h, int(n_steps), self._st(), C.byref(a), C.byref(b)), "mis_profile_step")
        self.frame += int(n_steps)
        return a.value, b.value

    def set_gather_mode(self, mode: int):
        """0: register-tiled cluster kernels (k_deform_c / k_force_c); 1: shared-memory cell tiles (k_deform_t / k_force_t);
        2: tiles for the deformation pass, cluster kernel for the force pass (include/mis.h)."""
        native.check(self.L.mis_set_gather_mode(self._h, int(mode), self._st()), "mis_set_gather_mode")

    def gather_info(self):
        out = (C.c_int * 8)()
        native.check(self.L.mis_get_gather_info(self._h, out), "mis_get_gather_info")
        keys = ("mode", "active_cells", "max_tile", "max_cell", "cap_deform", "cap_force", "list_blocks", "block_entries")
        return dict(zip(keys, [int(v) for v in out]))

    def kernel_names(self):
        """Names of the two gather kernels the step launches (what profile_step times; keys of profiles/traffic.json)."""
        mode = self.gather_info()["mode"]
        force = "k_force_t" if mode == 1 else "k_force_c"
        # scenes of >= 1500 clusters per SM run the force gather from per-SM cluster queues (k_force_p): same arithmetic, same
        # lists, another launch shape (enqueue_force in csrc/mis_api.cu; MIS_FORCE_PERSIST=0 switches it off)
        if force == "k_force_c" and not self.cfg.symmetric_pair and os.environ.get("MIS_FORCE_PERSIST", "1")[:1] != "0":
            nsm = torch.cuda.get_device_properties(self.device).multi_processor_count
            if self.n // (int(self.params.cluster_size) or 2) >= 1500 * nsm:
                force = "k_force_p"
        return {"k_deform": "k_deform_t" if mode else "k_deform_c", "k_force": force}

    @property
    def launch_count(self) -> int:
        return int(self.L.mis_launch_count(self._h))
