"""CPU suite, part 2: host-side logic, the C ABI's host-only entry points, exported symbols."""

import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import case_names, load, oracle_spec, params_array
from oracle import beam_oracle as bo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from continuum_robot_b200 import _lib

    return _lib


def test_library_exports_every_declared_symbol():
    """Every function declared in include/crb.h is exported by libcrb.so (no compute calls)."""
    hdr = open(os.path.join(ROOT, "include", "crb.h")).read()
    names = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(crb_\w+)\s*\(", hdr, flags=re.M))
    assert names >= {"crb_version", "crb_last_error", "crb_plan", "crb_assemble", "crb_rhs", "crb_forces", "crb_rk4", "crb_rk45",
                     "crb_dense_matrices", "crb_gain_fragments"}
    L = _lib()
    lib = L.load()
    for n in names:
        assert hasattr(lib, n), n
    assert set(L.EXPORTED_SYMBOLS) == names
    assert lib.crb_version() == L.CRB_VERSION
    assert C.sizeof(L.CrbPlan) == 4 * 10 + 8 * 2 + 4 * 3 * 256


def _plan(N, bc, mmax=0):
    L = _lib()
    p = L.CrbPlan()
    rc = L.load().crb_plan(N, bytes(bc), mmax, C.byref(p))
    return rc, p


@pytest.mark.parametrize("N,m,g", [(32, 4, 8), (64, 4, 16), (6, 3, 2), (10, 3, 4), (20, 3, 8), (1, 1, 1), (128, 4, 32)])
def test_plan_cantilever_shapes(N, m, g):
    rc, p = _plan(N, [1] + [0] * N)
    assert rc == 0
    assert (p.m, p.g, p.n0, p.p_act, p.n_free, p.contiguous, p.has_mask) == (m, g, 1, N, 3 * N, 1, 0)
    assert p.p == m * g >= N and p.levels == int(np.log2(g))
    red = np.ctypeslib.as_array(p.red_index)[: 3 * p.p]
    assert np.array_equal(red[: 3 * N], np.arange(3 * N)) and np.all(red[3 * N:] == -1)


@pytest.mark.parametrize("name", case_names(load("rhs_cases.npz")))
def test_plan_reduced_index_matches_reference_bc_reduction(name):
    """(slot, dof) -> reduced index equals the reference's sorted unconstrained-DOF list
    (euler_bernoulli_beam.py:256-258) for every BC pattern in the golden set."""
    g = load("rhs_cases.npz")
    spec = oracle_spec(g, name + "/")
    N = spec.n_elements
    bc = list(spec.bc) + [0]
    rc, p = _plan(N, bc)
    assert rc == 0
    unc = bo.unconstrained_dofs(spec)
    assert p.n_free == len(unc)
    red = np.ctypeslib.as_array(p.red_index)[: 3 * p.p]
    for r, full in enumerate(unc):
        node, d = divmod(int(full), 3)
        assert red[3 * (node - p.n0) + d] == r
    assert (red >= 0).sum() == len(unc)


def test_plan_errors_are_reported_not_raised():
    L = _lib()
    rc, _ = _plan(0, [0])
    assert rc < 0 and b"n_elements" in L.load().crb_last_error()
    rc, _ = _plan(3, [1, 1, 1, 1])
    assert rc < 0 and b"constrain all" in L.load().crb_last_error()
    rc, _ = _plan(2, [1, 7, 0])
    assert rc < 0 and b"boundary condition" in L.load().crb_last_error()
    rc, _ = _plan(200, [1] + [0] * 200)
    assert rc < 0 and b"limit" in L.load().crb_last_error()
    with pytest.raises(ValueError):
        L.check(rc)


@pytest.mark.parametrize("name", ["lin4", "pinned_root6", "interior_fixed6", "free6"])
def test_dense_matrices_match_reference(name):
    """crb_dense_matrices (host) vs the reference's get_mass_matrix / get_stiffness_matrix."""
    L = _lib()
    g = load("rhs_cases.npz")
    p = name + "/"
    par = np.ascontiguousarray(params_array(g, p))
    N = par.shape[0]
    bc = np.array(list(g[p + "bc"]) + [0], dtype=np.uint8)
    et = np.asarray(g[p + "elem_type"], dtype=np.uint8)
    rc, plan = _plan(N, bc)
    n = plan.n_free
    M = np.zeros((n, n))
    K = np.zeros((n, n))
    rc = L.load().crb_dense_matrices(C.byref(plan), par.ctypes.data_as(C.c_void_p), et.tobytes(), bc.tobytes(),
                                     M.ctypes.data_as(C.c_void_p), K.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert np.abs(M - g[p + "M"]).max() <= 1e-14 * np.abs(M).max()
    assert np.abs(K - g[p + "K"]).max() <= 1e-14 * np.abs(K).max()


def test_dense_stiffness_rejects_nonlinear():
    L = _lib()
    g = load("rhs_cases.npz")
    par = np.ascontiguousarray(params_array(g, "nl4/"))
    bc = np.array([1, 0, 0, 0, 0], dtype=np.uint8)
    rc, plan = _plan(4, bc)
    M = np.zeros((12, 12))
    K = np.zeros((12, 12))
    rc = L.load().crb_dense_matrices(C.byref(plan), par.ctypes.data_as(C.c_void_p), bytes([1] * 4), bc.tobytes(),
                                     M.ctypes.data_as(C.c_void_p), K.ctypes.data_as(C.c_void_p))
    assert rc < 0 and b"nonlinear" in L.load().crb_last_error()


# ---- functional-composition API semantics (reference tests/test_functional_composition.py) ----
class _MockForce:
    fused_kind = None

    def __init__(self, value, enabled=True):
        self.value, self.enabled, self.calls = value, enabled, 0

    def compute_forces(self, x, t):
        import torch

        self.calls += 1
        return torch.full((x.shape[0], x.shape[1] // 2), self.value, dtype=x.dtype)

    def is_enabled(self):
        return self.enabled


def test_force_registry_semantics():
    import torch

    from continuum_robot_b200 import ForceRegistry

    reg = ForceRegistry()
    a, b, off = _MockForce(1.0), _MockForce(2.5), _MockForce(9.0, enabled=False)
    reg.register(a)
    reg.register(b)
    reg.register(off)  # disabled instances are ignored at registration
    assert len(reg) == 2 and a in reg and off not in reg
    lst = reg.get_registered_forces()
    lst.clear()
    assert len(reg) == 2  # a copy was returned
    f = reg.create_aggregated_function()
    x = torch.zeros(3, 8, dtype=torch.float64)
    assert torch.equal(f(x, 0.0), torch.full((3, 4), 3.5, dtype=torch.float64))
    b.enabled = False  # toggling after creation takes effect on the next call
    assert torch.equal(f(x, 0.0), torch.full((3, 4), 1.0, dtype=torch.float64))
    assert reg.unregister(a) is True and reg.unregister(a) is False
    b.enabled = True
    reg.clear()
    assert len(reg) == 0 and torch.equal(f(x), torch.zeros(3, 4, dtype=torch.float64))


def test_input_registry_adds_deltas():
    import torch

    from continuum_robot_b200 import InputRegistry

    class H:
        def __init__(self, d, en=True):
            self.d, self.en = d, en

        def compute_input(self, x, r, t):
            return torch.full_like(r, self.d)

        def is_enabled(self):
            return self.en

    reg = InputRegistry()
    h1, h2 = H(1.0), H(0.25)
    reg.register(h1)
    reg.register(h2)
    u = torch.ones(2, 3, dtype=torch.float64)
    out = reg.create_aggregated_function()(torch.zeros(2, 6, dtype=torch.float64), u, 0.0)
    assert torch.equal(out, u + 1.25) and torch.equal(u, torch.ones(2, 3, dtype=torch.float64))
    assert len(reg.get_registered_handlers()) == 2 and h1 in reg


def test_force_params_and_properties_validation():
    from continuum_robot_b200 import ForceParams, Properties

    assert not ForceParams() and ForceParams(enable_gravity_effects=True)
    assert ForceParams(gravity_vector=[0, 0, 0], enable_gravity_effects=True).enable_gravity_effects is False
    with pytest.raises(ValueError, match="fluid_density must be positive"):
        ForceParams(enable_fluid_effects=True)
    with pytest.raises(ValueError, match="exactly 3 components"):
        ForceParams(gravity_vector=[0, 1])
    ok = dict(length=1.0, elastic_modulus=1.0, moment_inertia=1.0, density=1.0, cross_area=1.0, segment_id=0, element_type="Linear")
    assert Properties(**ok).get_element_type().value == "linear"
    for k, label in (("length", "Length"), ("density", "Density"), ("cross_area", "Cross area")):
        with pytest.raises(ValueError, match=f"{label} must be positive"):
            Properties(**{**ok, k: 0.0})
    with pytest.raises(ValueError, match="Invalid element type"):
        Properties(**{**ok, "element_type": "cubic"})


def test_lqr_host_synthesis_contract():
    """LinearQuadraticRegulator: shapes, block structure of A/B, stable closed loop, cached gain
    (the checks of the reference's tests/test_control.py)."""
    from continuum_robot_b200 import LinearQuadraticRegulator

    b = bo.BeamOracle(bo.BeamSpec.uniform(4))
    K, M = b.stiffness_matrix(), b.M
    n = K.shape[0]
    lqr = LinearQuadraticRegulator(K, M, np.eye(2 * n), np.eye(n))
    A, B = lqr.get_A(), lqr.get_B()
    assert np.array_equal(A[:n, n:], np.eye(n)) and not A[:n, :n].any() and not A[n:, n:].any() and not B[:n].any()
    G = lqr.compute_gain_matrix()
    assert G.shape == (n, 2 * n) and lqr.get_K() is G
    assert np.all(np.linalg.eigvals(A - B @ G).real < 0)
    with pytest.raises(ValueError, match="Q matrix dimension"):
        LinearQuadraticRegulator(K, M, np.eye(3), np.eye(n)).compute_gain_matrix()
    with pytest.raises(ValueError, match="same dimensions"):
        LinearQuadraticRegulator(K, np.eye(n + 1), np.eye(2 * n), np.eye(n))


def test_no_cpu_fallback():
    """The product refuses to run without a CUDA device instead of falling back to the oracle."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from continuum_robot_b200.dynamic_beam import BatchedDynamicEulerBernoulliBeam

    g = load("rhs_cases.npz")
    with pytest.raises((RuntimeError, AssertionError)):
        BatchedDynamicEulerBernoulliBeam({"params": params_array(g, "lin4/"), "type": ["linear"] * 4}, device="cpu")
    src = "".join(open(os.path.join(ROOT, "continuum_robot_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "continuum_robot_b200")) if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src


def test_gain_fragments_layout():
    """crb_gain_fragments: every gain entry appears exactly once at the (k-tile, n-tile, lane) slot
    the mma layout prescribes; plans without 4 lanes per member are rejected."""
    L = _lib()
    lib = L.load()
    rc, p = _plan(6, [1] + [0] * 6, 2)
    assert rc == 0 and (p.m, p.g) == (2, 4)
    n = p.n_free
    gain = np.arange(n * 2 * n, dtype=np.float64).reshape(n, 2 * n) + 1.0
    cnt = lib.crb_gain_fragments(C.byref(p), None, None)
    assert cnt == 12 * 3 * 32
    out = np.zeros(cnt)
    assert lib.crb_gain_fragments(C.byref(p), gain.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == cnt
    fr = out.reshape(12, 3, 32)
    assert sorted(fr[fr != 0].tolist()) == sorted(gain.ravel().tolist())
    red = np.ctypeslib.as_array(p.red_index)
    # k-tile 4 = own value #4 (slot 1, dof 1) of source lane 2; n-tile 1, element 0 = own dof #2 of dest lane 2
    lane = 2 + 4 * (2 * 2 + 0)
    assert fr[4, 1, lane] == gain[red[3 * (2 * 2) + 2], red[3 * (2 * 2) + 4]]
    assert fr[6 + 4, 1, lane] == gain[red[3 * (2 * 2) + 2], n + red[3 * (2 * 2) + 4]]
    assert not fr[:, :, 4 * 6:].any()  # destination lane 3 owns only phantom slots (6 active of 8)
    rc, p2 = _plan(6, [1] + [0] * 6, 0)
    assert lib.crb_gain_fragments(C.byref(p2), None, None) < 0 and b"4 lanes" in lib.crb_last_error()


@pytest.mark.parametrize("N,bc0,gravity,with_ref,imp,decoupled", [
    (6, 1, True, True, True, False), (3, 2, True, False, False, False), (8, 1, False, True, True, False), (1, 1, True, False, True, False),
    (6, 1, True, True, True, True), (5, 2, True, False, True, True), (8, 1, False, False, False, True)])
def test_shared_operator_blob_matches_numpy(N, bc0, gravity, with_ref, imp, decoupled):
    """crb_shared_operator (host only): de-tile the mma B fragments and compare W, c0, M^-1 e_k and the
    gravity operators with a NumPy restatement built from the oracle's M, K (dynamic_beam_model.py:256-272,
    full_state_linear.py:58, gravity_forces.py:97-146)."""
    L = _lib()
    lib = L.load()
    rng = np.random.default_rng(N)
    spec = bo.BeamSpec.uniform(N)
    spec.length = spec.length * (1 + 0.1 * rng.random(N))
    spec.density = spec.density * (1 + 0.2 * rng.random(N))
    spec.bc = np.array([bc0] + [0] * (N - 1))
    rc, plan = _plan(N, list(spec.bc) + [0])
    assert rc == 0
    n = plan.n_free
    par = np.ascontiguousarray(np.stack([spec.length, spec.elastic_modulus, spec.moment_inertia, spec.density, spec.cross_area,
                                         np.ones(N), np.ones(N)], axis=1))
    et = bytes(int(t) for t in spec.elem_type)
    bcb = bytes(int(b) for b in list(spec.bc) + [0])
    gain = rng.standard_normal((n, 2 * n))
    if decoupled:
        # a gain designed on a straight beam does not couple axial and bending DOFs (examples/lqr_control.py:46-84):
        # zero those blocks exactly, as SciPy's CARE solution has them (tests/golden/cfg5_samples.npz)
        axial = _axial_mask(spec.bc, N)
        cross = axial[:, None] != axial[None, :]
        gain[np.concatenate([cross, cross], axis=1)] = 0.0
    ref = rng.standard_normal(2 * n) if with_ref else None
    imp_dof = n - 2 if imp and n >= 2 else (0 if imp else -1)
    gx, gy = (0.0, -9.81) if decoupled else (1.5, -9.81)
    args = (C.byref(plan), par.ctypes.data_as(C.c_void_p), et, bcb, gain.ctypes.data_as(C.c_void_p),
            ref.ctypes.data_as(C.c_void_p) if with_ref else None, gx, gy, int(gravity), imp_dof)
    cnt = lib.crb_shared_operator(*args, None)
    assert cnt > 0
    blob = np.zeros(cnt)
    assert lib.crb_shared_operator(*args, blob.ctypes.data_as(C.c_void_p)) == cnt
    KQ, NT, GKP = int(blob[2]), int(blob[3]), int(blob[4])
    assert (int(blob[1]), KQ, NT, GKP, int(blob[7])) == (n, (n + 3) // 4, ((n + 3) // 4 + 1) // 2, (N + 3) // 4 if gravity else 0, cnt)
    tail = blob[cnt - 2 * KQ - 4:].view(np.int32)
    order, masks = tail[:4 * KQ], tail[4 * KQ:4 * KQ + 5]
    assert sorted(order[order >= 0].tolist()) == list(range(n))  # a permutation of the reduced DOFs plus padding
    o = 8
    wq = blob[o:o + 32 * KQ * NT].reshape(KQ, NT, 32); o += 32 * KQ * NT
    wv = blob[o:o + 32 * KQ * NT].reshape(KQ, NT, 32); o += 32 * KQ * NT
    pf = blob[o:o + 32 * KQ].reshape(KQ, 32); o += 32 * KQ
    gc = blob[o:o + 32 * GKP * NT].reshape(GKP, NT, 32); o += 32 * GKP * NT
    gs = blob[o:o + 32 * GKP * NT].reshape(GKP, NT, 32); o += 32 * GKP * NT
    c0i = blob[o:o + 4 * KQ]; o += 4 * KQ
    mii = blob[o:o + 4 * KQ]
    c0 = np.zeros(n); mi = np.zeros(n)
    c0[order[order >= 0]] = c0i[order >= 0]
    mi[order[order >= 0]] = mii[order >= 0]
    assert not c0i[order < 0].any() and not mii[order < 0].any()
    # de-tile: lane -> (k = lane % 4, column 2 jo + e = lane // 4); internal output index 4 (2 nt + e) + jo; internal index
    # r stands for reduced DOF order[r]
    W = np.zeros((n, 2 * n)); P = np.zeros((N, n)); Gc = np.zeros((n, N)); Gs = np.zeros((n, N))
    ordx = np.concatenate([order, -np.ones(8, dtype=np.int32)])
    for lane in range(32):
        k, ncol = lane % 4, lane // 4
        jo, e = ncol // 2, ncol % 2
        for i in range(KQ):
            c = ordx[4 * i + k]
            for nt in range(NT):
                out = ordx[4 * (2 * nt + e) + jo]
                if out >= 0 and c >= 0:
                    W[out, c], W[out, n + c] = wq[i, nt, lane], wv[i, nt, lane]
                else:
                    assert wq[i, nt, lane] == 0 and wv[i, nt, lane] == 0
            seg = 4 * e + jo
            if seg < N and c >= 0:
                P[seg, c] = pf[i, lane]
        for p in range(GKP):
            seg = 4 * p + k
            for nt in range(NT):
                out = ordx[4 * (2 * nt + e) + jo]
                if seg < N and out >= 0:
                    Gc[out, seg], Gs[out, seg] = gc[p, nt, lane], gs[p, nt, lane]
    # the tile masks name exactly the tiles that hold a nonzero
    for name, frag, mask in (("wq", wq, masks[0]), ("wv", wv, masks[1]), ("gc", gc, masks[3]), ("gs", gs, masks[4])):
        nz = frag.reshape(-1, 32).any(axis=1)
        assert [bool(mask >> b & 1) for b in range(len(nz))] == nz.tolist(), name
    assert [bool(masks[2] >> b & 1) for b in range(KQ)] == pf.any(axis=1).tolist()
    n_tiles = sum(bin(int(m) & 0xFFFFFFFF).count("1") for m in masks)
    if decoupled:
        assert n_tiles < 2 * KQ * NT + (KQ + 2 * GKP * NT if gravity else 0)  # the chosen order leaves whole tiles empty
        if (N, bc0, gravity) == (6, 1, True):
            assert n_tiles == 24  # config 5's shape: 16 (W) + 2 (P) + 4 (cos) + 2 (sin) of 47
    else:
        assert np.array_equal(order[:n], np.arange(n))  # a dense gain keeps the natural order
    orc = bo.BeamOracle(spec, bo.ForceSpec(gravity_vector=(gx, gy, 0.0), enable_gravity_effects=gravity))
    K = np.zeros((n, n))
    for j in range(n):
        e_j = np.zeros(n); e_j[j] = 1.0
        K[:, j] = orc.stiffness(e_j)
    Minv = np.linalg.inv(orc.M)
    Wref = np.concatenate([-Minv @ (K + gain[:, :n]), -Minv @ gain[:, n:]], axis=1)
    assert np.abs(W - Wref).max() <= 1e-11 * np.abs(Wref).max()
    if with_ref:
        assert np.abs(c0 - Minv @ (gain @ ref)).max() <= 1e-11 * np.abs(Minv @ (gain @ ref)).max()
    else:
        assert np.all(c0 == 0)
    if imp_dof >= 0:
        assert np.abs(mi - Minv[:, imp_dof]).max() <= 1e-11 * np.abs(Minv[:, imp_dof]).max()
    if gravity:
        # the operators reproduce the oracle's gravity force for random rotations: M^-1 f = Gc cos(P q) + Gs sin(P q)
        q = 0.3 * rng.standard_normal(n)
        x = np.concatenate([q, np.zeros(n)])
        want = Minv @ orc.gravity(x)
        got = Gc @ np.cos(P @ q) + Gs @ np.sin(P @ q)
        assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max()
    # argument errors are reported, not raised
    assert lib.crb_shared_operator(C.byref(plan), None, et, bcb, None, None, gx, gy, 0, -1, blob.ctypes.data_as(C.c_void_p)) < 0
    assert lib.crb_shared_operator(C.byref(plan), par.ctypes.data_as(C.c_void_p), et, bcb, None, None, gx, gy, 0, n, None) < 0


def _axial_mask(bc, N):
    """True for the axial (u) entries of the BC-reduced DOF vector (euler_bernoulli_beam.py:221-298: node i keeps
    [u, w, phi] when free, [phi] when pinned, nothing when fixed)."""
    out = []
    for b in list(bc) + [0]:
        if b == 0:
            out += [True, False, False]
        elif b == 2:
            out += [False]
    return np.array(out)


@pytest.mark.parametrize("N", [3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("gravity", [False, True])
def test_shared_operator_of_the_lqr_example_fits_the_compiled_tile_pattern(N, gravity):
    """The straight FIXED-root beam of examples/lqr_control.py:26-84 (gain without axial / bending coupling, gravity
    along y): the tile masks crb_shared_operator writes are a subset of the masks the rollout kernel's specialised
    path is compiled for (crb_shared_sparse_masks), i.e. that path is the one that runs -- and it skips tiles."""
    lib = _lib().load()
    spec = bo.BeamSpec.uniform(N)
    rc, plan = _plan(N, [1] + [0] * N)
    assert rc == 0
    n = plan.n_free
    par = np.ascontiguousarray(np.stack([spec.length, spec.elastic_modulus, spec.moment_inertia, spec.density, spec.cross_area,
                                         np.ones(N), np.ones(N)], axis=1))
    et = bytes(int(t) for t in spec.elem_type)
    bcb = bytes([1] + [0] * N)
    gain = np.random.default_rng(N).standard_normal((n, 2 * n))
    axial = _axial_mask(spec.bc, N)
    cross = axial[:, None] != axial[None, :]
    gain[np.concatenate([cross, cross], axis=1)] = 0.0
    args = (C.byref(plan), par.ctypes.data_as(C.c_void_p), et, bcb, gain.ctypes.data_as(C.c_void_p), None, 0.0, -9.81, int(gravity), n - 2)
    cnt = lib.crb_shared_operator(*args, None)
    blob = np.zeros(cnt)
    assert lib.crb_shared_operator(*args, blob.ctypes.data_as(C.c_void_p)) == cnt
    KQ, NT, GKP = int(blob[2]), int(blob[3]), int(blob[4])
    masks = blob[cnt - 2 * KQ - 4:].view(np.uint32)[4 * KQ:4 * KQ + 5]
    compiled = np.zeros(5, dtype=np.uint32)
    assert lib.crb_shared_sparse_masks(KQ, GKP, compiled.ctypes.data_as(C.c_void_p)) == 0
    assert compiled[0] != 0xFFFFFFFF, "no specialised path for this shape"
    assert not (masks & ~compiled).any(), (masks, compiled)
    full = 2 * KQ * NT + (KQ + 2 * GKP * NT if gravity else 0)
    used = sum(bin(int(m)).count("1") for m in compiled)
    assert used < full
    assert lib.crb_shared_sparse_masks(0, 0, compiled.ctypes.data_as(C.c_void_p)) < 0


def test_shared_operator_size_limits():
    lib = _lib().load()
    rc, plan = _plan(9, [1] + [0] * 9)  # n_free = 27 > 24
    assert rc == 0
    assert lib.crb_shared_operator(C.byref(plan), None, None, None, None, None, 0.0, 0.0, 0, -1, None) < 0
    assert b"n_free" in lib.crb_last_error()


def test_system_slice_offsets_per_member_pointers():
    """crb_system_slice (host only): per-member arrays are offset, shared ones are not."""
    L = _lib()
    lib = L.load()
    rc, plan = _plan(32, [1] + [0] * 32)
    assert rc == 0
    s = L.CrbSystem()
    s.n_members = 1000
    s.mass_shared, s.stiff_shared, s.force_shared = 1, 0, 0
    base = 1 << 20
    s.mfac, s.kcoef, s.drag, s.grav, s.seg_half_mass = base, 2 * base, 3 * base, 4 * base, 5 * base
    s.u_const, s.f_ext, s.imp_amp, s.elem_type, s.red_index = 6 * base, 7 * base, 8 * base, 9 * base, 10 * base
    out = L.CrbSystem()
    assert lib.crb_system_slice(C.byref(plan), C.byref(s), 100, 50, C.byref(out)) == 0
    P, n, N = plan.p, plan.n_free, plan.n_elements
    assert out.n_members == 50 and out.mfac == base and out.elem_type == 9 * base and out.red_index == 10 * base
    assert out.kcoef == 2 * base + 8 * 100 * P * 4
    assert out.drag == 3 * base + 8 * 100 * P and out.grav == 4 * base + 8 * 100 * P * 2
    assert out.seg_half_mass == 5 * base + 8 * 100 * N
    assert out.u_const == 6 * base + 8 * 100 * n and out.f_ext == 7 * base + 8 * 100 * n and out.imp_amp == 8 * base + 8 * 100
    s.mass_shared = 0
    assert lib.crb_system_slice(C.byref(plan), C.byref(s), 7, 1, C.byref(out)) == 0
    assert out.mfac == base + 8 * 7 * plan.mfac_doubles
    assert lib.crb_system_slice(C.byref(plan), C.byref(s), 990, 20, C.byref(out)) < 0
    assert b"outside" in lib.crb_last_error()


def test_midpoint_entry_points_validate_before_touching_the_device():
    """crb_assemble_shifted / crb_midpoint report bad arguments through the return code (no CUDA call yet)."""
    L = _lib()
    lib = L.load()
    rc, plan = _plan(4, [1, 0, 0, 0, 0])
    assert rc == 0
    par = np.ones((1, 4, 7))
    args = (C.byref(plan), par.ctypes.data_as(C.c_void_p), 1, bytes([0, 1, 0, 0]), bytes([1, 0, 0, 0, 0]), 1)
    assert lib.crb_assemble_shifted(*args, 1e-8, C.c_void_p(16), None) < 0
    assert b"not linear" in lib.crb_last_error()
    lin = (C.byref(plan), par.ctypes.data_as(C.c_void_p), 1, bytes([0, 0, 0, 0]), bytes([1, 0, 0, 0, 0]))
    assert lib.crb_assemble_shifted(*lin, 1, -1.0, C.c_void_p(16), None) < 0 and b"shift" in lib.crb_last_error()
    assert lib.crb_assemble_shifted(*lin, 3, 1e-8, C.c_void_p(16), None) < 0 and b"n_sets" in lib.crb_last_error()
    assert lib.crb_assemble_shifted(*lin, 1, 1e-8, None, None) < 0
    s = L.CrbSystem()
    s.n_members = 4
    assert lib.crb_midpoint(C.byref(plan), C.byref(s), C.c_void_p(16), 1, C.c_void_p(16), 0.0, 1e-4, 1, None, 0, None) < 0
    assert b"not assembled" in lib.crb_last_error()
    s.mfac = s.kcoef = s.elem_type = s.red_index = 16
    s.all_linear = 1
    s.grav_mode = 1
    s.grav = 16
    assert lib.crb_midpoint(C.byref(plan), C.byref(s), C.c_void_p(16), 1, C.c_void_p(16), 0.0, 1e-4, 1, None, 0, None) < 0
    assert b"all-linear beam without drag / gravity" in lib.crb_last_error()
    s.grav_mode, s.grav = 0, None
    assert lib.crb_midpoint(C.byref(plan), C.byref(s), C.c_void_p(16), 1, C.c_void_p(16), 0.0, 0.0, 1, None, 0, None) < 0
    assert b"positive" in lib.crb_last_error()
    assert lib.crb_midpoint(C.byref(plan), C.byref(s), C.c_void_p(16), 1, C.c_void_p(16), 0.0, 1e-4, 0, None, 0, None) == 0  # nothing to do


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference from oracle/_ref where oracle/build_ref.py installed it, else the
    CPU port; no GPU needed): one JSON line with the driver-contract keys; under a multi-rank launch only rank 0
    prints."""
    import json
    import subprocess
    import sys

    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "12", "--warmup", "3"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    have_ref = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "continuum_robot"))
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] and "workload" in d["config"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_lqr_entry_points_validate_arguments_on_the_host():
    """crb_lqr_* / crb_dense_matrices_batched reject bad arguments before touching the device (no compute calls)."""
    L = _lib()
    lib = L.load()
    err = lambda: lib.crb_last_error().decode()
    need = C.c_size_t(0)
    assert lib.crb_lqr_workspace_bytes(97, 8, C.byref(need)) == -3 and "exceed the limit" in err()  # CRB_LQR_MAX_N
    assert lib.crb_lqr_workspace_bytes(0, 8, C.byref(need)) == -1
    assert lib.crb_lqr_workspace_bytes(18, 8, None) == -1
    assert lib.crb_lqr_gains(18, 4, None, 0, None, 0, None, None, 1, None, None, None, None, None, 0, None) == -1
    assert "null argument" in err()
    one = C.c_void_p(8)  # never dereferenced: the size checks come first
    assert lib.crb_lqr_gains(100, 4, one, 0, one, 0, one, one, 1, one, None, None, one, one, 0, None) == -3
    rc, p = _plan(4, [1, 0, 0, 0, 0])
    assert rc == 0
    et_nl, bc = bytes([0, 1, 0, 0]), bytes([1, 0, 0, 0, 0])
    assert lib.crb_dense_matrices_batched(C.byref(p), None, 1, et_nl, bc, None, None, None) == -1
    assert lib.crb_dense_matrices_batched(C.byref(p), one, 1, et_nl, bc, one, one, None) == -1
    assert "nonlinear segments" in err()  # euler_bernoulli_beam.py:422-456
    assert lib.crb_dense_matrices_batched(C.byref(p), one, 0, bytes(4), bc, one, None, None) == -1
    assert lib.crb_dense_matrices_batched(C.byref(p), one, 1, bytes(4), bytes([2, 0, 0, 0, 0]), one, None, None) == -1
    assert "mismatch" in err()
    assert lib.crb_member_operators(18, 4, None, 0, None, 0, None, None, None, None, None) == -1
    assert lib.crb_member_operators(33, 4, one, 0, one, 0, one, None, one, one, None) == -3 and "one lane per DOF" in err()
    # the ctypes mirror has the compiled layout (crb_abi_sizes), fields in header order
    pb, sb = C.c_int32(), C.c_int32()
    assert lib.crb_abi_sizes(C.byref(pb), C.byref(sb)) == 0
    assert (pb.value, sb.value) == (C.sizeof(L.CrbPlan), C.sizeof(L.CrbSystem))
    assert C.sizeof(L.CrbSystem) % 8 == 0 and L.CrbSystem.out_sel_inv.offset == C.sizeof(L.CrbSystem) - 24
    assert L.CrbSystem.member_order.offset == C.sizeof(L.CrbSystem) - 8
    assert L.CrbSystem.member_op.offset == L.CrbSystem.gain_stride.offset + 8 == L.CrbSystem.u_sin_amp.offset - 8


def test_header_is_plain_c_and_the_library_serves_a_c_client(tmp_path):
    """include/crb.h compiles as C99 (no C++ / torch types in the boundary) and a plain-C program linked against
    libcrb.so gets the same answers from the host-only entry points as the ctypes binding."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    L = _lib()
    lib = L.load()
    libdir = os.path.dirname(L.LIB_PATH)
    exe = str(tmp_path / "host_smoke")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "host_smoke.c"), "-o", exe,
                           "-L" + libdir, "-l:libcrb.so", "-Wl,-rpath," + libdir, "-Wl,--allow-shlib-undefined"])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    vals = dict(line.split(" ", 1) for line in out.strip().splitlines())
    assert int(vals["version"]) == L.CRB_VERSION
    rc, p = _plan(4, [1, 0, 0, 0, 0])
    assert vals["plan"] == f"n_free {p.n_free} m {p.m} g {p.g} p {p.p} contiguous {p.contiguous}"
    par = np.tile(np.array([0.25, 75e9, 4.908738521234052e-10, 6450.0, 7.853981633974483e-05, 0.007853981633974483, 0.82]), (4, 1))
    M, K = np.zeros((12, 12)), np.zeros((12, 12))
    assert lib.crb_dense_matrices(C.byref(p), par.ctypes.data_as(C.c_void_p), bytes(4), bytes([1, 0, 0, 0, 0]),
                                  M.ctypes.data_as(C.c_void_p), K.ctypes.data_as(C.c_void_p)) == 0
    serial = lambda A: float(sum(float(A[i, i]) for i in range(12)))  # same left-to-right order as the C loop
    assert float(vals["trace_M"]) == serial(M) and float(vals["trace_K"]) == serial(K)
    assert float(vals["asym_K"]) == 0.0
    assert vals["bad_bc"].startswith("rc -1") and int(vals["bad_bc"].split()[-1]) > 10


def test_rk45_pilot_launch_policy(monkeypatch):
    """Host logic of solve_ensemble's pilot launch (integrate._pilot_attempts): on for ensembles of more than one and
    at most 64 waves of resident warps, off otherwise; an explicit request or CRB_RK45_PILOT wins; never the whole
    attempt budget."""
    from types import SimpleNamespace

    from continuum_robot_b200.integrate import _pilot_attempts

    monkeypatch.delenv("CRB_RK45_PILOT", raising=False)
    beam = SimpleNamespace(_plan=SimpleNamespace(g=32), device="cuda:0")  # one member per warp
    slots = 8 * 148
    assert _pilot_attempts(beam, slots, None, 10**6, sm_count=148) == 0          # one wave: everything starts at once
    assert _pilot_attempts(beam, slots + 1, None, 10**6, sm_count=148) == 8
    assert _pilot_attempts(beam, 4096, None, 10**6, sm_count=148) == 8           # BASELINE config 4
    assert _pilot_attempts(beam, 64 * slots, None, 10**6, sm_count=148) == 8
    assert _pilot_attempts(beam, 64 * slots + 1, None, 10**6, sm_count=148) == 0  # the tail is short against the launch
    assert _pilot_attempts(beam, 4096, None, 8, sm_count=148) == 0                # budget too small to split
    half = SimpleNamespace(_plan=SimpleNamespace(g=16), device="cuda:0")          # two members per warp
    assert _pilot_attempts(half, 2 * slots, None, 10**6, sm_count=148) == 0
    assert _pilot_attempts(half, 2 * slots + 2, None, 10**6, sm_count=148) == 8
    assert _pilot_attempts(beam, 10, 5, 10**6, sm_count=148) == 5                 # explicit request
    assert _pilot_attempts(beam, 4096, 0, 10**6, sm_count=148) == 0
    assert _pilot_attempts(beam, 4096, 50, 20, sm_count=148) == 19                # at least one attempt is left for the main launch
    monkeypatch.setenv("CRB_RK45_PILOT", "3")
    assert _pilot_attempts(beam, 10, None, 10**6, sm_count=148) == 3
    assert _pilot_attempts(beam, 10, 7, 10**6, sm_count=148) == 7                 # the argument wins over the environment


def test_integration_md_stub_mirrors_the_struct():
    """The ctypes stub shown in INTEGRATION.md lists the fields of crb_system_t in the order of the product's mirror."""
    import os
    import re

    L = _lib()
    txt = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "INTEGRATION.md")).read()
    blk = txt[txt.index("class CrbSystem(C.Structure)"):]
    blk = blk[:blk.index("]\n", blk.index("member_order"))]
    assert re.findall(r'\("(\w+)",', blk) == [f[0] for f in L.CrbSystem._fields_]
