"""smb_chain: op-chain fusion.  The fused result must be what the
reference's separate operators produce one after another -- each SMArray operator
(include/SMArray.h:217-305) materialises a rounded intermediate, so the oracle is the chain of
oracle operators on the same inputs: bit-exact for + - * / (float, double, int32 wrap), the
array_scalar_op lane / scalar-tail split for int32 pow, the stated ULP bound for float pow."""
import ctypes

import numpy as np
import pytest

import oracle
import simplemath_b200 as smb
from conftest import assert_same_bits


def oracle_chain(orc, first, steps):
    """The unfused sequence on the CPU oracle.  Returns the last intermediate too (for pow ULPs)."""
    acc, prev = np.ascontiguousarray(first), None
    for op, leaf in steps:
        swap = op.startswith("r") and op[1:] in smb.OPS
        op = op[1:] if swap else op
        prev = acc
        if isinstance(leaf, np.ndarray):
            acc = orc.binary(op, leaf, acc) if swap else orc.binary(op, acc, leaf)
        elif swap:
            acc = orc.binary(op, np.full((1,) * acc.ndim, leaf, acc.dtype), acc)
        else:
            acc = orc.array_scalar(op, acc, leaf)
    return acc, prev


# ---------------------------------------------------------------- no GPU needed
def test_chain_argument_checks_need_no_device():
    lib = smb.lib()
    shape = smb._u64arr([4])
    steps = smb.chain_steps(smb.F32, [(None, False, 1.0), ("add", False, 2.0)], [4])
    out = np.zeros(4, np.float32)
    assert lib.smb_chain(smb.F32, steps, 0, shape, 1, 4, out.ctypes.data, None) == 1          # no steps
    assert lib.smb_chain(smb.F32, steps, 9, shape, 1, 4, out.ctypes.data, None) == 1          # too many
    assert lib.smb_chain(7, steps, 2, shape, 1, 4, out.ctypes.data, None) == 1                # dtype
    assert lib.smb_chain(smb.F32, steps, 2, shape, 1, 5, out.ctypes.data, None) == 1          # n != prod(shape)
    bad = smb.chain_steps(smb.F32, [(None, False, 1.0), ("pow", False, (out.ctypes.data, [1]))], [4])
    assert lib.smb_chain(smb.F32, bad, 2, shape, 1, 4, out.ctypes.data, None) == 1            # array exponent
    assert b"constant exponent" in lib.smb_last_error()
    if smb.device_count() == 0:   # valid arguments: fails loudly without a GPU, never computes on the CPU
        assert lib.smb_chain(smb.F32, steps, 2, shape, 1, 4, out.ctypes.data, None) == 2
        assert b"no CPU fallback" in lib.smb_last_error()


def test_chain_step_struct_matches_header():
    hdr = open(smb.HERE + "/../include/smb200.h").read()
    assert "#define SMB_CHAIN_MAX 8" in hdr and smb.CHAIN_MAX == 8
    assert ctypes.sizeof(smb.ChainStep) == 4 + 4 + 8 + 8 * smb.MAX_NDIM + 8


def test_chain_planner_coalescing_keeps_every_offset():
    """smb_plan_chain (host only): dims merged across ALL leaves; the offset of every flat index of every
    leaf is unchanged, constants coalesce with anything, and the known configs collapse as expected."""
    rng = np.random.default_rng(5)

    def offsets(shape, strides):
        idx = np.indices(shape).reshape(len(shape), -1)
        return (idx * np.array(strides, dtype=np.int64)[:, None]).sum(0)

    for _ in range(200):
        nd = int(rng.integers(1, 7))
        shape = [int(x) for x in rng.integers(1, 5, size=nd)]
        leaves, tables = [], []
        for s in range(int(rng.integers(1, 9))):
            if s and rng.random() < 0.25:
                leaves.append(("add", False, 1.0))
                tables.append(None)
                continue
            dims = [d if rng.random() < 0.7 else 1 for d in shape]
            st = smb.row_major_strides(dims)
            if rng.random() < 0.3:
                st = [x * 2 for x in st]                         # a view with padded parent strides
            st = [0 if d == 1 and r > 1 else x for d, r, x in zip(dims, shape, st)]
            leaves.append((None if s == 0 else "add", False, (4096, st)))
            tables.append(st)
        if tables[0] is None:
            continue
        vec, pshape, pstrides = smb.plan_chain(smb.F32, leaves, shape)
        assert int(np.prod(pshape)) == int(np.prod(shape))
        for st, pst in zip(tables, pstrides):
            if st is None:
                assert all(x == 0 for x in pst)
            else:
                assert np.array_equal(offsets(shape, st), offsets(pshape, pst))
        assert vec == all(pst[-1] <= 1 for pst in pstrides)
    # a dense 3-leaf chain of rank 4 is one dim; config C2 as a chain stays 2-D; C4's pattern 3-D
    dense = smb.row_major_strides([4, 5, 6, 8])
    assert smb.plan_chain(smb.F32, [(None, False, (4096, dense)), ("add", False, (4096, dense)), ("mul", False, 2.0)], [4, 5, 6, 8])[1] == [960]
    assert smb.plan_chain(smb.F32, [(None, False, (4096, [4096, 1])), ("add", False, (4096, [0, 1]))], [4096, 4096])[1] == [4096, 4096]
    vec, sh, st = smb.plan_chain(smb.I32, [(None, False, (4096, [1024, 0, 1])), ("mul", False, (4096, [0, 1024, 1]))], [512, 512, 1024])
    assert vec and sh == [512, 512, 1024] and st == [[1024, 0, 1], [0, 1024, 1]]


def load_golden_chains(name="golden_chain_v1.npz"):
    """tests/golden/golden_chain_v1.npz: chains evaluated operator by operator on the UNMODIFIED
    reference (oracle/make_golden_chain.py).  Yields (first, steps, out); a "fix" entry replaces the
    accumulator before a `leaf / acc` step (the generator made it a safe divisor), so such a chain
    is checked in two pieces.  golden_chain_v2.npz (same layout, same generator): int32 sm::pow on an
    intermediate that a LATER leaf broadcasts up."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    i = 0
    while f"k{i:03d}_first" in z:
        p = f"k{i:03d}_"
        ops = [str(o) for o in z[p + "ops"]]
        first, steps = z[p + "first"], []
        for s, op in enumerate(ops):
            if p + f"fix{s}" in z:            # restart from the injected intermediate
                first, steps = z[p + f"fix{s}"], []
            leaf = z[p + f"leaf{s}"]
            steps.append((op, leaf if leaf.ndim else leaf.item()))
        yield i, first, steps, z[p + "out"]
        i += 1


def test_golden_chain_fixtures_agree_with_the_oracle(orc):
    """The oracle's operator-by-operator chain reproduces the reference's (pins oracle_chain below)."""
    n = 0
    for i, first, steps, out in load_golden_chains():
        want, _ = oracle_chain(orc, first, steps)
        assert_same_bits(np.asarray(want).reshape(out.shape), out, f"golden chain {i}")
        n += 1
    assert n == 38


def test_golden_int32_pow_on_broadcast_intermediates_agrees_with_the_oracle(orc):
    """golden_chain_v2: the reference's lane / scalar-tail split of int32 sm::pow follows the flat index
    of the N-element temporary it is applied to, not of the M*N-element chain result."""
    n = 0
    for i, first, steps, out in load_golden_chains("golden_chain_v2.npz"):
        want, _ = oracle_chain(orc, first, steps)
        assert_same_bits(np.asarray(want).reshape(out.shape), out, f"golden chain v2 {i}")
        n += 1
    assert n == 115


# ------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_chain_matches_reference_golden_fixtures():
    """Fused smb_chain against outputs of the unmodified reference's separate operators, bit for bit."""
    n = 0
    for i, first, steps, out in load_golden_chains():
        assert_same_bits(smb.chain(first, *steps).reshape(out.shape), out, f"golden chain {i}: {[s[0] for s in steps]}")
        n += 1
    assert n == 38


@pytest.mark.gpu
def test_chain_int32_pow_on_broadcast_intermediates_matches_reference_golden():
    """golden_chain_v2 through the fused smb_chain: `pow(row, e) + B`, `pow(row + 1, e) * col`, ... must
    split lanes / tail by the intermediate's own flat index (the round-1 kernel used the output's)."""
    n = 0
    for i, first, steps, out in load_golden_chains("golden_chain_v2.npz"):
        assert_same_bits(smb.chain(first, *steps).reshape(out.shape), out, f"golden chain v2 {i}: {[s[0] for s in steps]}")
        n += 1
    assert n == 115


pytestmark_gpu = pytest.mark.gpu


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
def test_chain_bit_exact_against_unfused_oracle_sequence(orc, dt):
    rng = np.random.default_rng(71)

    def rnd(shape):
        if dt == np.int32:
            return rng.integers(-2**31, 2**31, size=shape).astype(np.int32)
        return (rng.standard_normal(shape) * 100).astype(dt)

    def nz(shape):  # divisors
        if dt == np.int32:
            return rng.integers(1, 1000, size=shape).astype(np.int32) * rng.choice(np.array([-1, 1], np.int32), size=shape)
        return (rng.uniform(0.5, 3, size=shape) * rng.choice([-1, 1], size=shape)).astype(dt)

    c3 = 3 if dt == np.int32 else 2.5
    a, b, c, d = rnd((64, 256)), rnd((64, 256)), nz((64, 256)), rnd((64, 256))
    cases = [
        (a, [("add", b)]),
        (a, [("add", b), ("mul", c)]),
        (a, [("add", b), ("mul", c), ("sub", d)]),
        (a, [("sub", b), ("div", c), ("add", c3), ("rsub", d), ("mul", a)]),
        (a, [("mul", c3), ("radd", b), ("rdiv", c3)] if dt != np.int32 else [("mul", c3), ("radd", b), ("div", c)]),
        # broadcast leaves: a row, a column, a rank-1 row, constants in between
        (a, [("add", b[:1]), ("mul", c[:, :1]), ("sub", d[0]), ("add", c3)]),
        (a[:, :1], [("add", b[:1]), ("mul", c)]),                       # the first leaf is the broadcast one
        (rnd((8, 1, 32)), [("mul", rnd((1, 16, 32))), ("add", rnd((8, 16, 1))), ("sub", rnd((32,)))]),  # 3-D
        (rnd((5, 3, 2, 4, 2, 8)), [("add", rnd((5, 1, 2, 1, 2, 8))), ("mul", rnd((1, 3, 1, 4, 1, 1)))]),  # rank 6
        (rnd((1000,)), [("add", rnd((1000,))), ("mul", rnd((1000,))), ("sub", rnd((1000,))), ("add", rnd((1000,))),
                        ("mul", rnd((1000,))), ("sub", rnd((1000,))), ("add", rnd((1000,)))]),           # 8 leaves
        (rnd((7, 13)), [("add", rnd((7, 13))), ("mul", rnd((13,)))]),   # odd inner length -> scalar variant
        (a.T, [("add", b.T), ("mul", c.T)]),                            # transposed leaves -> scalar variant
        (a[:, 1:65], [("add", b[:, 3:67]), ("sub", d[:, :64])]),        # misaligned interior pointers
    ]
    for first, steps in cases:
        want, _ = oracle_chain(orc, first, steps)
        got = smb.chain(first, *steps)
        assert got.shape == want.shape
        assert_same_bits(got, want, f"{np.dtype(dt).name} chain {[s[0] for s in steps]} {first.shape}")
    smb.chain(a, ("add", b), ("mul", c))
    assert smb.last_kernel() == "k_chain<vec16>"
    smb.chain(a.T, ("add", b.T))
    assert smb.last_kernel() == "k_chain<scalar>"


@pytest.mark.gpu
def test_chain_int_pow_lane_and_tail_semantics(orc):
    rng = np.random.default_rng(72)
    for n in (5, 8, 1003, 4096):
        a = rng.integers(-6, 7, size=n).astype(np.int32)
        b = rng.integers(-3, 4, size=n).astype(np.int32)
        for e in (3, 2, 20, -2, 0, 31):
            steps = [("add", b), ("pow", e), ("sub", 1)]
            want, _ = oracle_chain(orc, a, steps)
            assert_same_bits(smb.chain(a, *steps), want, f"int pow chain n={n} e={e}")


@pytest.mark.gpu
@pytest.mark.parametrize("dt,bound", [(np.float32, 1.0), (np.float64, 1.0)])
def test_chain_float_pow_within_ulp_bound(orc, dt, bound):
    rng = np.random.default_rng(73)
    a = rng.uniform(0.01, 50, size=(128, 512)).astype(dt)
    b = rng.uniform(0.01, 50, size=(1, 512)).astype(dt)
    for y in (2.0, 2.5, -0.75, 7.0, 11.5, -30.0):
        steps = [("add", b), ("pow", y)]
        _, base = oracle_chain(orc, a, steps)          # base = a + b, bit-exact
        got = smb.chain(a, *steps)
        if dt == np.float32:
            # the table-driven core with staged tables -- except y = 2, which takes its exact form (acc * acc) like the eager operator
            assert smb.last_kernel() == ("k_chain<vec16>" if y == 2.0 else "k_chain<vec16,pow>"), smb.last_kernel()
        assert_same_bits(smb.chain(a, ("add", b)), base, "pow base")
        if dt == np.float32:
            err = oracle.ulp_error_f32(got.ravel(), orc.pow_ref_f32(base.ravel(), float(np.float32(y))))
        else:
            hi, lo = orc.pow_ref_f64(base.ravel(), y)
            err = oracle.ulp_error_f64(got.ravel(), hi, lo)
        assert err.max() <= bound, (y, err.max())
        # and a step after the pow consumes the rounded power
        got2 = smb.chain(a, ("add", b), ("pow", y), ("mul", a))
        assert_same_bits(got2, orc.binary("mul", got, a), "step after pow")
    # negative bases: integer exponents keep / drop the sign, non-integer ones give NaN; specials fall back
    x = np.concatenate([rng.uniform(-50, 50, 4090), [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-42]]).astype(dt)
    z = np.zeros_like(x)
    for y in (3.0, 4.0, -3.0, 2.5, 0.5):
        got = smb.chain(x, ("add", z), ("pow", y))
        if dt == np.float32:
            ref = orc.pow_ref_f32(orc.binary("add", x, z), float(np.float32(y)))   # (-0) + 0 is +0
            with np.errstate(all="ignore"):
                want = ref.astype(np.float32)
            fin = np.isfinite(want) & (want != 0)
            assert np.array_equal(np.isnan(got), np.isnan(want)), y
            assert np.array_equal(got[~fin & ~np.isnan(want)], want[~fin & ~np.isnan(want)]), y
            assert np.array_equal(np.signbit(got[~np.isnan(want)]), np.signbit(want[~np.isnan(want)])), y
            assert oracle.ulp_error_f32(got[fin], ref[fin]).max() <= bound, y


@pytest.mark.gpu
def test_chain_wide_index_ranges_and_device_pointers(orc):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(74)
    a = rng.standard_normal((33, 40)).astype(np.float32)
    b = rng.standard_normal((1, 40)).astype(np.float32)
    c = rng.standard_normal((33, 1)).astype(np.float32)
    steps = [("mul", b), ("add", c), ("sub", 0.5)]
    want, _ = oracle_chain(orc, a, steps)
    smb.set_option(smb.OPT_FORCE_WIDE_INDEX, 1)
    try:
        assert_same_bits(smb.chain(a, *steps), want, "wide index")
        assert smb.last_kernel() == "k_chain<vec16,wide>"
    finally:
        smb.set_option(smb.OPT_FORCE_WIDE_INDEX, 0)
    # device-resident leaves, flat sub-ranges (the multi-GPU shard unit), asynchronous on a stream
    ta, tb, tc = (torch.from_numpy(x).cuda() for x in (a, b, c))
    n = a.size
    leaves = [(None, False, (ta.data_ptr(), [40, 1])), ("mul", False, (tb.data_ptr(), [0, 1])),
              ("add", False, (tc.data_ptr(), [1, 0])), ("sub", False, 0.5)]
    sp = torch.cuda.current_stream().cuda_stream
    for lo, cnt in ((0, n), (40, 400), (4, 1316), (13, 77)):
        out = torch.zeros(cnt, dtype=torch.float32, device="cuda")
        smb.chain_ptr(smb.F32, leaves, [33, 40], out.data_ptr(), stream=sp, lin_range=(lo, cnt))
        torch.cuda.synchronize()
        assert_same_bits(out.cpu().numpy(), want.ravel()[lo:lo + cnt], f"range {lo}+{cnt}")
    l0 = smb.launch_count()
    out = torch.zeros(n, dtype=torch.float32, device="cuda")
    smb.chain_ptr(smb.F32, leaves, [33, 40], out.data_ptr(), stream=sp)
    torch.cuda.synchronize()
    assert smb.launch_count() == l0 + 1      # ONE kernel for the whole chain


@pytest.mark.gpu
def test_chain_full_size_properties():
    """At a size the oracle would not finish quickly: (x + y) - y == x wherever the sum is exact,
    and fused == the library's own unfused operators, bit for bit."""
    torch = pytest.importorskip("torch")
    n = 1 << 26
    sp = torch.cuda.current_stream().cuda_stream
    x = torch.randint(-2**20, 2**20, (n,), device="cuda").to(torch.float32)
    y = torch.randint(-2**20, 2**20, (n,), device="cuda").to(torch.float32)
    z = torch.rand(n, device="cuda") + 0.5
    out, t1, t2 = (torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(3))
    leaves = [(None, False, (x.data_ptr(), [1])), ("add", False, (y.data_ptr(), [1])), ("sub", False, (y.data_ptr(), [1]))]
    smb.chain_ptr(smb.F32, leaves, [n], out.data_ptr(), stream=sp)
    torch.cuda.synchronize()
    assert torch.equal(out, x)
    leaves = [(None, False, (x.data_ptr(), [1])), ("add", False, (y.data_ptr(), [1])), ("div", False, (z.data_ptr(), [1])),
              ("mul", False, 3.0)]
    smb.chain_ptr(smb.F32, leaves, [n], out.data_ptr(), stream=sp)
    smb.contiguous_ptr(smb.OP_ADD, smb.F32, x.data_ptr(), y.data_ptr(), t1.data_ptr(), n, sp)
    smb.contiguous_ptr(smb.OP_DIV, smb.F32, t1.data_ptr(), z.data_ptr(), t2.data_ptr(), n, sp)
    smb.array_scalar_ptr(smb.OP_MUL, smb.F32, t2.data_ptr(), 3.0, n, t1.data_ptr(), sp)
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int32), t1.view(torch.int32))


@pytest.mark.gpu
def test_chain_randomised_shapes_ranks_and_leaf_kinds(orc):
    """Seeded fuzz over everything the planner and the rank-specialised kernels branch on: result rank
    1..6, per-leaf broadcast patterns (any subset of dims of size 1, lower rank), sliced / transposed views,
    constants, swapped operands, chain lengths 1..8, the three dtypes, 32- and 64-bit index forms."""
    rng = np.random.default_rng(20261020)
    ops_all = ["add", "sub", "mul", "div", "rsub", "rdiv"]
    kernels = set()
    for case in range(160):
        dt = [np.float32, np.float64, np.int32][case % 3]
        rank = int(rng.integers(1, 7))
        shape = [int(rng.choice([1, 2, 3, 4, 5, 8, 12, 16])) for _ in range(rank)]
        if rng.random() < 0.5:
            shape[-1] = int(rng.choice([4, 8, 16, 32, 64]))      # often a vectorisable inner dim

        def leaf_array(divisor):
            keep = [d if rng.random() < 0.7 else 1 for d in shape]
            drop = int(rng.integers(0, rank))                     # lower rank: leading dims missing
            lshape = keep[drop:] if rng.random() < 0.3 and drop < rank else keep
            kind = rng.random()
            full = [max(2 * d, 2) if kind < 0.25 else d for d in lshape]
            if dt == np.int32:
                base = rng.integers(1, 50, size=full) if divisor else rng.integers(-2**31, 2**31, size=full, dtype=np.int64)
                base = (base * (rng.choice([-1, 1], size=full) if divisor else 1)).astype(np.int32)
            else:
                base = (rng.uniform(0.5, 2, size=full) * rng.choice([-1, 1], size=full)).astype(dt) if divisor \
                    else (rng.standard_normal(full) * 100).astype(dt)
            if kind < 0.25:                                        # an interior slice of a larger block
                sl = tuple(slice(1, 1 + d) if f > d else slice(None) for d, f in zip(lshape, full))
                return base[sl]
            if kind < 0.4 and len(lshape) >= 2:                    # a transposed view of the right shape
                return np.ascontiguousarray(base.swapaxes(-1, -2)).swapaxes(-1, -2)
            return base

        first = leaf_array(False)
        steps = []
        for _ in range(int(rng.integers(0, 8))):
            op = str(rng.choice(ops_all))
            if rng.random() < 0.25 and not op.startswith("r"):
                c = int(rng.integers(1, 9)) if dt == np.int32 else float(rng.uniform(0.5, 4))
                steps.append((op, c))
            else:
                steps.append((op, leaf_array(divisor=op == "div")))
        # `leaf / acc` needs acc != 0 (and != -1 for INT_MIN): make the chain safe by construction
        safe = []
        for op, leaf in steps:
            if op == "rdiv":
                safe.append(("mul", 0 if dt == np.int32 else 0.0))
                safe.append(("add", 3 if dt == np.int32 else 1.5))
                if len(safe) + 1 > smb.CHAIN_MAX - 1:
                    break
            safe.append((op, leaf))
            if len(safe) >= smb.CHAIN_MAX - 1:
                break
        steps = safe[: smb.CHAIN_MAX - 1]
        try:
            want, _ = oracle_chain(orc, first, steps)
        except RuntimeError:
            continue                                               # incompatible shapes: both sides refuse
        if np.asarray(want).ndim > smb.MAX_NDIM:
            continue
        wide = case % 5 == 0
        if wide:
            smb.set_option(smb.OPT_FORCE_WIDE_INDEX, 1)
        try:
            got = smb.chain(first, *steps)
        finally:
            if wide:
                smb.set_option(smb.OPT_FORCE_WIDE_INDEX, 0)
        kernels.add(smb.last_kernel())
        assert_same_bits(got, np.asarray(want).reshape(got.shape), f"fuzz case {case}: {np.dtype(dt).name} {shape} {[s[0] for s in steps]}")
    assert {"k_chain<vec16>", "k_chain<scalar>", "k_chain<vec16,wide>", "k_chain<scalar,wide>"} <= kernels, kernels


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
def test_fuzz_chains_over_random_views_of_device_arrays(orc, dt):
    """~120 random chains per dtype: 2-7 steps, each leaf a random strided / permuted / broadcast view of a DEVICE-resident
    parent (interior pointers, odd bases) or a constant, operators + - * / in both operand orders -- fused result against
    the oracle's operator-by-operator sequence, bit for bit."""
    import torch
    from test_gpu_parity import _random_view
    tdt = {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32}[dt]
    rng = np.random.default_rng(20261103 + np.dtype(dt).itemsize + (5 if dt == np.int32 else 0))
    dims = [1, 2, 3, 4, 8, 12, 16, 33, 64, 100]
    code = smb.dtype_code(np.dtype(dt))
    es = np.dtype(dt).itemsize
    kernels = {}
    for case in range(120):
        nd = int(rng.integers(1, 5))
        shape = [int(rng.choice(dims)) for _ in range(nd)]
        while int(np.prod(shape)) > 300_000:
            shape[int(rng.integers(0, nd))] = 2
        friendly = case % 2 == 0                    # every other case qualifies for the vector kernel
        if friendly:
            shape[-1] = int(rng.choice([4, 8, 16, 64, 100]))
        nsteps = int(rng.integers(2, 8))
        keep = []                                   # device parents stay alive until the launch is checked
        steps_np, leaves = [], []
        for s in range(nsteps):
            op = "add" if s == 0 else str(rng.choice(["add", "sub", "mul", "div", "rsub", "add", "mul"]))
            if s > 0 and rng.random() < 0.2:        # a constant
                v = int(rng.integers(1, 9)) if dt == np.int32 else float(np.float32(rng.uniform(0.5, 4.0)))
                steps_np.append((op, v))
                leaves.append((op[1:] if op == "rsub" else op, op == "rsub", v))
                continue
            want = [1 if (s > 0 and rng.random() < 0.3) else d for d in shape]
            drop = int(rng.integers(0, nd)) if (s > 0 and rng.random() < 0.2) else 0
            parent, view = _random_view(rng, dt, want[drop:], op == "div", friendly)
            dev = torch.from_numpy(parent.copy()).cuda()
            keep.append(dev)
            ptr = dev.data_ptr() + (view.__array_interface__["data"][0] - parent.__array_interface__["data"][0])
            st = [0] * drop + [x // es for x in view.strides]
            dd = [1] * drop + list(view.shape)
            st = [0 if d == 1 and r > 1 else x for d, r, x in zip(dd, shape, st)]
            steps_np.append((op, view))
            leaves.append((None if s == 0 else (op[1:] if op == "rsub" else op), op == "rsub", (ptr, st)))
        first = steps_np[0][1]
        want_out, _ = oracle_chain(orc, np.broadcast_to(first, shape) if first.shape != tuple(shape) else first, steps_np[1:])
        out = torch.zeros(int(np.prod(shape)) + 4, dtype=tdt, device="cuda")
        torch.cuda.synchronize()
        smb.chain_ptr(code, leaves, shape, out.data_ptr())
        k = smb.last_kernel()
        kernels[k] = kernels.get(k, 0) + 1
        host = out.cpu().numpy()
        assert_same_bits(host[:-4].reshape(shape), np.asarray(want_out).reshape(shape), f"chain case {case}: {[o for o, _ in steps_np]} {shape} [{k}]")
        assert not host[-4:].any(), f"chain case {case}: wrote past the result [{k}]"
    assert any(k.startswith("k_chain<vec16") for k in kernels) and any(k.startswith("k_chain<scalar") for k in kernels), kernels
