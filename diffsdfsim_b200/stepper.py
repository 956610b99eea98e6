"""Host side of the device-resident step loop (csrc/dsdf_steploop.cu): buffers, burst launches, tape, reverse sweep.

``World3D.step`` (lcp_physics/physics/world.py:119-139, 241-379) becomes ONE autograd node per step:

* forward  -- ``DeviceStepper.run`` launches rounds (prep -> solve -> move -> find_contacts -> commit) in bursts; the
  per-world accept / reject / halve / remaining-time / time-of-contact decisions are taken on the device; the host reads
  one 64-byte control block per burst (normally once per step).  Accepted sub-steps are recorded on a TAPE
  (``dsdf_step_slot``): slot k holds the k-th accepted sub-step of every world in this step.
* backward -- ``_StepFn.backward`` walks the tape slots in reverse and chains the hand-written VJP kernels
  (time-of-contact -> contact geometry -> integrator -> implicit LCP backward) with per-world masks; no autograd graph
  is recorded per round and rejected attempts leave no trace, exactly as in the reference (world.py:344-356 restores
  the state before retrying).
"""
import ctypes

import torch

from . import _lib
from .contacts import ContactSet

F64 = torch.float64
U8 = torch.uint8
I32 = torch.int32
BASE_TOL = 1e-6      # lcp_physics/physics/utils.py:43 (World.H.backward, world.py:204)

MAX_SLOTS = 16
STEP_CAPK, STEP_MAXC, STEP_DYN_SMEM, STEP_TAPE, STEP_MAX_ROUNDS = 1, 2, 4, 8, 16
CT_ABORT, CT_NACT, CT_MAXCOUNT, CT_ROUNDS, CT_MAXNSUB, CT_ANYTOC, CT_LCPSTAT = 0, 1, 3, 4, 5, 8, 9
LCP_FACTOR_FAIL, LCP_INACCURATE, LCP_TOO_LARGE = 4, 8, 16

_SLOT_FIELDS = ['p_in', 'v_in', 'x', 'new_v', 'p_try', 'dt_raw', 'dt_used', 'lam', 's', 'toc_flag_in', 'toc_now',
                'toc_mask', 'count', 'body', 'face', 'abc', 'geo']
_INT_FIELDS = ['W', 'nb', 'neq', 'maxc', 'fric_dirs', 'capK', 'npairs', 'depth', 'spec_threshold', 'vcap', 'n_slots',
               'max_iter', 'max_rounds', 'strict', 'toc_enabled', 'fixed_dt', 'detach_b2']
_DBL_FIELDS = ['world_dt', 'eps', 'tol', 'fd_eps', 'body_eps']
_PTR_FIELDS = ['geom', 'pairs', 'eq_rows', 'mass', 'Ibody', 'fric', 'rest', 'f', 'shape',
               'p', 'v', 't', 'dt_try', 'end_t', 'last_dt', 'active', 'toc_flag', 'had', 'nsub', 'attempts',
               'count', 'status', 'body', 'face', 'abc', 'geo',
               'vmap', 'vidx', 'dt_raw_v', 'dt_used_v',
               'x_v', 'new_v_v', 'nu_v', 'lam_v', 's_v', 'p_try_v', 'lcp_status_v', 'iters_v',
               'count_v', 'status_v', 'body_v', 'face_v', 'abc_v', 'geo_v', 'ctrl']


class StepSlot(ctypes.Structure):
    """dsdf_step_slot (include/dsdf_b200.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in _SLOT_FIELDS]


class StepArgs(ctypes.Structure):
    """dsdf_step_args (include/dsdf_b200.h): every field 8 bytes wide, same order."""
    _fields_ = ([(n, ctypes.c_int64) for n in _INT_FIELDS] + [(n, ctypes.c_double) for n in _DBL_FIELDS]
                + [(n, ctypes.c_void_p) for n in _PTR_FIELDS] + [('slots', StepSlot * MAX_SLOTS)])


def _ptr(t):
    return None if t is None else _lib.ptr(t)


class TapeSlot:
    """Device buffers of one dsdf_step_slot."""

    def __init__(self, W, nb, maxc, per, dev):
        e = lambda *s: torch.empty(*s, dtype=F64, device=dev)
        self.p_in, self.v_in, self.p_try = e(W, nb, 7), e(W, nb, 6), e(W, nb, 7)
        self.x, self.new_v = e(W, 6 * nb), e(W, nb, 6)
        self.dt_raw, self.dt_used = torch.zeros(W, dtype=F64, device=dev), torch.zeros(W, dtype=F64, device=dev)
        self.lam, self.s = e(W, maxc * per), e(W, maxc * per)
        self.toc_flag_in = torch.zeros(W, dtype=U8, device=dev)
        self.toc_now = torch.zeros(W, dtype=U8, device=dev)
        self.toc_mask = torch.zeros(W, maxc, dtype=U8, device=dev)
        self.cs = ContactSet(W, maxc, dev)           # zero-filled: count = 0 for worlds that never reach this slot

    def regrown(self, maxc, per):
        """The same records with room for ``maxc`` contacts per world."""
        W, nb, dev, old = self.p_in.shape[0], self.p_in.shape[1], self.p_in.device, self.toc_mask.shape[1]
        n = TapeSlot.__new__(TapeSlot)
        n.__dict__.update(self.__dict__)
        n.lam, n.s = (torch.empty(W, maxc * per, dtype=F64, device=dev) for _ in range(2))
        n.lam[:, :old * per], n.s[:, :old * per] = self.lam, self.s
        n.toc_mask = torch.zeros(W, maxc, dtype=U8, device=dev)
        n.toc_mask[:, :old] = self.toc_mask
        n.cs = self.cs.resized(maxc)
        return n

    def fill(self, c_slot):
        for k in ('p_in', 'v_in', 'x', 'new_v', 'p_try', 'dt_raw', 'dt_used', 'lam', 's', 'toc_flag_in', 'toc_now',
                  'toc_mask'):
            setattr(c_slot, k, _ptr(getattr(self, k)))
        for k in ('count', 'body', 'face', 'abc', 'geo'):
            setattr(c_slot, k, _ptr(getattr(self.cs, k)))


class Tape:
    """Everything one step leaves behind for its reverse sweep."""
    __slots__ = ('start', 'slots', 'nsub', 'maxsub', 'any_toc', 'maxc', 'C', 'rounds', 'p_out', 'v_out', 'geo_out',
                 'last_dt_out', 'had', 'final', 'f', 'syncs')


class DeviceStepper:
    """Owns the per-world work buffers of the step loop of one ``World3D`` and launches its rounds."""
    INITIAL_SLOTS = 2

    def __init__(self, world):
        _lib.lib()
        self.W, self.nb, self.dev = world.W, world.nb, world.device
        W, dev = self.W, self.dev
        self.depth = world.SPEC_DEPTH if world.speculate else 1
        self.spec_threshold = W // 8 if (world.speculate and W >= 64) else 0
        self.vcap = max(W, self.depth * self.spec_threshold)
        self.ctrl = torch.zeros(16, dtype=I32, device=dev)
        self.ctrl_host = torch.zeros(16, dtype=I32).pin_memory()
        self.dt_try = torch.zeros(W, dtype=F64, device=dev)
        self.end_t = torch.zeros(W, dtype=F64, device=dev)
        self.active = torch.zeros(W, dtype=U8, device=dev)
        self.nsub = torch.zeros(W, dtype=I32, device=dev)
        self.vidx = torch.zeros(W, dtype=I32, device=dev)
        self.maxc = None
        self.last_rounds = 1
        self.max_count = 0
        self.launches = 0
        self._alloc_virtual(world)

    def _alloc_virtual(self, world):
        V, nb, dev = self.vcap, self.nb, self.dev
        per = 2 + world.fric_dirs
        self.maxc = world.maxc
        e = lambda *s: torch.empty(*s, dtype=F64, device=dev)
        self.vmap = torch.zeros(V, dtype=I32, device=dev)
        self.dt_raw_v, self.dt_used_v = e(V), e(V)
        self.x_v, self.new_v_v, self.p_try_v = e(V, 6 * nb), e(V, 6 * nb), e(V, nb * 7)
        self.nu_v = e(V, max(world.num_constraints, 1))
        self.lam_v, self.s_v = e(V, self.maxc * per), e(V, self.maxc * per)
        self.lcp_status_v = torch.zeros(V, dtype=I32, device=dev)
        self.iters_v = torch.zeros(V, dtype=I32, device=dev)
        self.cs_v = ContactSet(V, self.maxc, dev)

    # ------------------------------------------------------------------------------------------------------------
    def _args(self, world, fixed_dt, st, tape, f):
        a = StepArgs()
        a.W, a.nb, a.neq, a.maxc, a.fric_dirs = self.W, self.nb, world.num_constraints, world.maxc, world.fric_dirs
        a.capK, a.npairs = world.detector.capK, world.detector.npairs
        a.depth, a.spec_threshold, a.vcap, a.n_slots = self.depth, self.spec_threshold, self.vcap, len(tape.slots)
        a.max_iter, a.max_rounds = world.engine.max_iter, world.max_rounds_per_step
        a.strict, a.toc_enabled = int(world.strict_no_pen), int(world.time_of_contact_diff)
        a.fixed_dt, a.detach_b2 = int(fixed_dt), int(world.detach_contact_b2)
        a.world_dt, a.eps, a.tol, a.fd_eps, a.body_eps = float(world.dt), world.eps, world.tol, 1e-3, world.body_eps
        a.geom, a.pairs, a.eq_rows = world.table.ptr(), _ptr(world.detector.pairs), _ptr(world.eq_rows)
        a.mass, a.Ibody, a.fric, a.rest = _ptr(st['mass']), _ptr(st['Ibody']), _ptr(st['fric']), _ptr(st['rest'])
        a.f, a.shape = _ptr(f), _ptr(world.shape)
        a.p, a.v, a.t, a.dt_try, a.end_t = _ptr(st['p']), _ptr(st['v']), _ptr(st['t']), _ptr(self.dt_try), _ptr(self.end_t)
        a.last_dt, a.active, a.toc_flag, a.had = _ptr(st['last_dt']), _ptr(self.active), _ptr(st['toc_flag']), _ptr(st['had'])
        a.nsub, a.attempts = _ptr(self.nsub), _ptr(world.stats['attempts'])
        cur = st['cur']
        a.count, a.status, a.body, a.face, a.abc, a.geo = (_ptr(cur.count), _ptr(cur.status), _ptr(cur.body),
                                                           _ptr(cur.face), _ptr(cur.abc), _ptr(cur.geo))
        a.vmap, a.vidx, a.dt_raw_v, a.dt_used_v = _ptr(self.vmap), _ptr(self.vidx), _ptr(self.dt_raw_v), _ptr(self.dt_used_v)
        a.x_v, a.new_v_v, a.nu_v, a.lam_v, a.s_v = (_ptr(self.x_v), _ptr(self.new_v_v), _ptr(self.nu_v), _ptr(self.lam_v),
                                                    _ptr(self.s_v))
        a.p_try_v, a.lcp_status_v, a.iters_v = _ptr(self.p_try_v), _ptr(self.lcp_status_v), _ptr(self.iters_v)
        v = self.cs_v
        a.count_v, a.status_v, a.body_v, a.face_v, a.abc_v, a.geo_v = (_ptr(v.count), _ptr(v.status), _ptr(v.body),
                                                                       _ptr(v.face), _ptr(v.abc), _ptr(v.geo))
        a.ctrl = _ptr(self.ctrl)
        for k, s in enumerate(tape.slots):
            s.fill(a.slots[k])
        return a

    def _smem_contacts(self, world):
        """Contacts the dynamics kernel sizes its shared memory for: the largest accepted count so far + slack."""
        c = (self.max_count + 2 + 3) // 4 * 4
        return max(4, min(c, world.maxc, 64))

    def _read_ctrl(self):
        self.ctrl_host.copy_(self.ctrl, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.ctrl_host.tolist()

    def run(self, world, fixed_dt, p, v, last_dt, mass, Ibody, fric, rest, f):
        """One World3D.step for all worlds.  Returns the Tape (which carries the new state)."""
        W, nb, dev = self.W, self.nb, self.dev
        per = 2 + world.fric_dirs
        if self.maxc != world.maxc:
            self._alloc_virtual(world)
        self.max_count = max(self.max_count, int(world.max_nc))
        tape = Tape()
        tape.start = world.contact_set                      # immutable from here on (the loop works on a clone)
        tape.slots = [TapeSlot(W, nb, world.maxc, per, dev) for _ in range(self.INITIAL_SLOTS)]
        tape.f = f
        st = dict(p=p.detach().clone(), v=v.detach().clone(), t=world.t.clone(), last_dt=last_dt.detach().clone(),
                  toc_flag=world.toc_flag.clone(), had=torch.zeros(W, dtype=U8, device=dev),
                  cur=world.contact_set.clone(), mass=mass, Ibody=Ibody, fric=fric, rest=rest)
        st['cur'].pre_ids = st['cur'].pre_cnt = None
        stream = _lib.stream()
        args = self._args(world, fixed_dt, st, tape, f)
        _lib.check(_lib.call('dsdf_step_begin', ctypes.byref(args), stream), 'dsdf_step_begin')
        burst = min(max(2, self.last_rounds + 1), 24)
        syncs = 0
        while True:
            rc = _lib.call('dsdf_step_rounds', ctypes.byref(args), burst, self._smem_contacts(world), stream)
            _lib.check(rc, 'dsdf_step_rounds')
            self.launches += 5 * burst
            c = self._read_ctrl()
            syncs += 1
            self.max_count = max(self.max_count, c[CT_MAXCOUNT])
            ab = c[CT_ABORT]
            if ab & STEP_MAX_ROUNDS:
                stuck = self.active.nonzero().flatten().tolist()[:8]
                raise RuntimeError('step did not complete in %d attempts: worlds %s keep penetrating (dt < %.3g); '
                                   'use strict_no_penetration=False or a smaller dt'
                                   % (world.max_rounds_per_step, stuck, float(self.dt_try.min())))
            if ab:
                self._grow(world, ab, tape, st, per)
                args = self._args(world, fixed_dt, st, tape, f)
                _lib.check(_lib.call('dsdf_step_resume', ctypes.byref(args), stream), 'dsdf_step_resume')
                continue
            if c[CT_NACT] == 0:
                break
            burst = min(2 * burst, 32)
        self.last_rounds = c[CT_ROUNDS]
        ls = c[CT_LCPSTAT]
        if ls & LCP_TOO_LARGE:
            raise _lib.DsdfLibraryError('dynamics kernel: a world had more contacts than its shared memory holds')
        world.engine.last_status_bits = ls
        if ls & LCP_INACCURATE and getattr(world.engine, 'verbose', -1) >= 0:
            print('qpth warning: Returning an inaccurate and potentially incorrect solution.')       # batch.py:165,229
        tape.nsub, tape.maxsub, tape.any_toc = self.nsub.clone(), c[CT_MAXNSUB], bool(c[CT_ANYTOC])
        tape.maxc, tape.C, tape.rounds, tape.syncs = world.maxc, self._smem_contacts(world), c[CT_ROUNDS], syncs
        tape.p_out, tape.v_out, tape.last_dt_out, tape.had = st['p'], st['v'], st['last_dt'], st['had'].bool()
        tape.final = st['cur']
        tape.geo_out = st['cur'].geo.clone()
        world.t, world.toc_flag = st['t'], st['toc_flag']
        del tape.slots[max(tape.maxsub, 0):]
        return tape

    def _grow(self, world, bits, tape, st, per):
        """Enlarge what the paused worlds ran out of (the worlds themselves are untouched and still active)."""
        if bits & STEP_CAPK:
            capK = world.detector.capK
            if capK >= world.MAX_CAPK:
                raise RuntimeError('contact candidate capacity exceeded at the kernel limit capK=%d (mesh too fine '
                                   'for the one-CTA-per-world contact kernel)' % capK)
            world._set_capacity(min(world.MAX_CAPK, 2 * capK), world.maxc)
        if bits & STEP_MAXC:
            maxc = world.maxc
            if maxc >= world.MAX_MAXC:
                raise RuntimeError('more than %d contacts in one world: beyond the dynamics kernel limit' % maxc)
            new = min(world.MAX_MAXC, 2 * maxc)
            world._set_capacity(world.detector.capK, new)
            st['cur'] = st['cur'].resized(new)
            tape.start = tape.start.resized(new)
            tape.slots = [s.regrown(new, per) for s in tape.slots]
            self._alloc_virtual(world)
        if bits & STEP_TAPE:
            if len(tape.slots) >= MAX_SLOTS:
                raise RuntimeError('a world accepted more than %d sub-steps inside one step' % MAX_SLOTS)
            n = min(MAX_SLOTS, 2 * len(tape.slots))
            tape.slots += [TapeSlot(self.W, self.nb, world.maxc, per, self.dev) for _ in range(n - len(tape.slots))]
        if bits & STEP_DYN_SMEM and self.max_count + 2 > 64:
            raise RuntimeError('more than 62 contacts in one world: beyond the one-warp dynamics kernel')


# ---------------------------------------------------------------------------------------------------- raw VJP calls
def _integrate_bwd(p, v, dt, active, g):
    W, nb = p.shape[0], p.shape[1]
    gp, gv = torch.empty_like(p), torch.empty_like(v)
    gdt = torch.empty(W, nb, dtype=F64, device=p.device)
    rc = _lib.call('dsdf_integrate_backward', _ptr(p), _ptr(v), _ptr(dt), _ptr(active), W, nb, _ptr(g), _ptr(gp),
                   _ptr(gv), _ptr(gdt), _lib.stream())
    _lib.check(rc, 'dsdf_integrate_backward')
    return gp, gv, gdt.sum(1)


def _toc_bwd(dt, toc_mask, body, p, v, geo, f, mass, gh):
    W, nb, maxc = p.shape[0], p.shape[1], geo.shape[1]
    g_dt, gp, gv = torch.empty_like(dt), torch.empty_like(p), torch.empty_like(v)
    ggeo, gf, gm = torch.empty_like(geo), torch.empty_like(f), torch.empty_like(mass)
    rc = _lib.call('dsdf_toc_backward', W, nb, maxc, _ptr(dt), _ptr(toc_mask), _ptr(body), _ptr(p), _ptr(v), _ptr(geo),
                   _ptr(f), _ptr(mass), _ptr(gh), BASE_TOL, _ptr(g_dt), _ptr(gp), _ptr(gv), _ptr(ggeo), _ptr(gf), _ptr(gm),
                   _lib.stream())
    _lib.check(rc, 'dsdf_toc_backward')
    return g_dt, gp, gv, ggeo, gf, gm


def _geometry_bwd(table, p, shape, cs, ggeo, fd_eps, detach_b2):
    W, nb = p.shape[0], p.shape[1]
    gp = torch.empty_like(p)
    rc = _lib.call('dsdf_contact_geometry_backward', table.ptr(), _ptr(p), _ptr(shape), W, nb, fd_eps, int(detach_b2),
                   cs.maxc, _ptr(cs.count), _ptr(cs.body), _ptr(cs.face), _ptr(cs.abc), _ptr(ggeo), _ptr(gp), _lib.stream())
    _lib.check(rc, 'dsdf_contact_geometry_backward')
    return gp


def _dyn_bwd(cfg, p, v, mass, Ibody, fric, rest, f, dt, active, cs, x, lam, s, gnv):
    W, nb = p.shape[0], p.shape[1]
    geo = cs.geo
    gp, gv = torch.empty_like(p), torch.empty_like(v)
    gmass, gI = torch.empty_like(mass), torch.empty_like(Ibody)
    gfric, grest, gf = torch.empty_like(fric), torch.empty_like(rest), torch.empty_like(f)
    gdt, ggeo = torch.empty(W, dtype=F64, device=p.device), torch.empty_like(geo)
    rc = _lib.call('dsdf_dynamics_solve_backward', _ptr(p), _ptr(v), _ptr(mass), _ptr(Ibody), _ptr(fric), _ptr(rest),
                   _ptr(f), _ptr(dt), _ptr(active), _ptr(cs.count), _ptr(cs.body), _ptr(geo), _ptr(cfg['eq_rows']), W, nb,
                   cfg['neq'], cs.maxc, cfg['C'], cfg['fric_dirs'], int(cfg['stop_contact_grad']),
                   int(cfg['stop_friction_grad']), _ptr(x), _ptr(lam), _ptr(s), _ptr(gnv), _ptr(gp), _ptr(gv), _ptr(gmass),
                   _ptr(gI), _ptr(gfric), _ptr(grest), _ptr(gf), _ptr(gdt), _ptr(ggeo), _lib.stream())
    _lib.check(rc, 'dsdf_dynamics_solve_backward')
    return gp, gv, gmass, gI, gfric, grest, gf, gdt, ggeo


class _StepFn(torch.autograd.Function):
    """One ``World3D.step`` of all worlds: (p, v, contact geometry, last_dt, parameters, forces) -> the same after the step.

    Per accepted sub-step (world.py:249-341), with cs_in the contacts found at the end of the previous sub-step:
        dt_    = toc_flag ? -last_dt + (last_dt.detach() + dt) : dt
        new_v  = solve(p, v, parameters, f, dt_, geo_in)                    engines.py:31-83
        p_try  = move(p, new_v, dt_)                                        bodies.py:488-496
        geo    = contact geometry(p_try)                                    contacts.py:262-264
        new contact:  dt_h = H(dt_, p_try, new_v, geo, f, mass);  p' = move(p, new_v, dt_h);  last_dt' = dt_h
        otherwise:    p' = p_try
    """

    @staticmethod
    def forward(ctx, world, fixed_dt, p, v, geo, last_dt, mass, Ibody, fric, rest, f):
        p, v, mass, Ibody, fric, rest, f = [t.contiguous() for t in (p, v, mass, Ibody, fric, rest, f)]
        tape = world._stepper.run(world, fixed_dt, p, v, last_dt, mass, Ibody, fric, rest, f)
        ctx.tape = tape
        ctx.params = (mass, Ibody, fric, rest, f)
        ctx.geo_in_maxc = geo.shape[1]
        ctx.cfg = dict(eq_rows=world.eq_rows, neq=world.num_constraints, C=tape.C, fric_dirs=world.fric_dirs,
                       stop_contact_grad=world.stop_contact_grad, stop_friction_grad=world.stop_friction_grad,
                       table=world.table, shape=world.shape, detach_b2=world.detach_contact_b2)
        world._last_tape = tape
        ctx.mark_non_differentiable(tape.had)
        return tape.p_out, tape.v_out, tape.geo_out, tape.last_dt_out, tape.had

    @staticmethod
    def backward(ctx, gp, gv, ggeo, glast, _ghad):
        tape, cfg = ctx.tape, ctx.cfg
        mass, Ibody, fric, rest, f = ctx.params
        W = mass.shape[0]
        maxc = tape.maxc
        gp, gv, glast = gp.contiguous(), gv.contiguous(), glast.contiguous()
        if ggeo.shape[1] != maxc:
            ggeo = torch.cat([ggeo, ggeo.new_zeros(W, maxc - ggeo.shape[1], 10)], 1)
        ggeo = ggeo.contiguous()
        gm, gI = torch.zeros_like(mass), torch.zeros_like(Ibody)
        gfr, gre, gf = torch.zeros_like(fric), torch.zeros_like(rest), torch.zeros_like(f)
        zero_w = torch.zeros(W, dtype=F64, device=mass.device)
        for k in range(tape.maxsub - 1, -1, -1):
            S = tape.slots[k]
            cin = tape.start if k == 0 else tape.slots[k - 1].cs
            m = tape.nsub > k
            m_u8 = m.to(U8)
            m3 = m[:, None, None]
            gp_in_acc = gnv_acc = None
            gdt_acc, glast_pass, ggeo_tot, gptry = zero_w, glast, ggeo, gp
            if tape.any_toc:
                # new contacts: p' = move(p, new_v, H(dt_)); H = dsdf_toc_backward (world.py:141-237, 275-341)
                tn_u8 = S.toc_now * m_u8
                tn = tn_u8.bool()
                gpA, gvA, gdtA = _integrate_bwd(S.p_in, S.new_v, S.dt_used, tn_u8, gp)
                gh = torch.where(tn, gdtA + glast, zero_w)
                g_dt_B, gp_B, gv_B, ggeo_B, gf_B, gm_B = _toc_bwd(S.dt_used, S.toc_mask, S.cs.body, S.p_try, S.new_v,
                                                                   S.cs.geo, f, mass, gh)
                tn3 = tn[:, None, None]
                gp_in_acc = torch.where(tn3, gpA, torch.zeros_like(gpA))
                gnv_acc = torch.where(tn3, gvA + gv_B, torch.zeros_like(gvA))
                gptry = torch.where(tn3, gp_B, gp)
                ggeo_tot = ggeo + torch.where(tn3, ggeo_B, torch.zeros_like(ggeo_B))
                gdt_acc = torch.where(tn, g_dt_B, zero_w)
                glast_pass = torch.where(tn, zero_w, glast)
                gf = gf + torch.where(tn3, gf_B, torch.zeros_like(gf_B))
                gm = gm + torch.where(tn[:, None], gm_B, torch.zeros_like(gm_B))
            gptry = gptry + _geometry_bwd(cfg['table'], S.p_try, cfg['shape'], S.cs, ggeo_tot.contiguous(), 1e-3,
                                          cfg['detach_b2'])
            gp2, gv2, gdt2 = _integrate_bwd(S.p_in, S.new_v, S.dt_used, m_u8, gptry.contiguous())
            gnv = gv + gv2 if gnv_acc is None else gv + gv2 + gnv_acc      # masked worlds: gv2 = 0, gv passes through
            gp3, gv3, gm3, gI3, gfr3, gre3, gf3, gdt3, ggeo3 = _dyn_bwd(cfg, S.p_in, S.v_in, mass, Ibody, fric, rest, f,
                                                                       S.dt_used, m_u8, cin, S.x, S.lam, S.s,
                                                                       gnv.contiguous())
            gp_new = gp2 + gp3 if gp_in_acc is None else gp2 + gp3 + gp_in_acc
            gdt_tot = gdt_acc + gdt2 + gdt3
            glast = torch.where(m, glast_pass + torch.where(S.toc_flag_in.bool(), -gdt_tot, zero_w), glast)
            gp = torch.where(m3, gp_new, gp)
            gv = gv3
            ggeo = torch.where(m3, ggeo3, ggeo)
            gm, gI, gfr, gre, gf = gm + gm3, gI + gI3, gfr + gfr3, gre + gre3, gf + gf3
        if ggeo.shape[1] != ctx.geo_in_maxc:
            ggeo = ggeo[:, :ctx.geo_in_maxc]
        return None, None, gp, gv, ggeo, glast, gm, gI, gfr, gre, gf
