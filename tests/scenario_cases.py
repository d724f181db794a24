"""The scenario cases shared by the golden generator (tests/golden/make_scenario_golden.py: reference on CPU, frozen
into tests/golden/scenarios.npz) and by the GPU parity tests (tests/test_gpu_live_reference.py: package under test
vs the frozen outputs AND vs the live reference on the same GPU).

(case id, scenario function in tests/scenarios.py, kwargs, comparison mode, kwargs override for the reference side)
modes: 'exact' bit-for-bit; 'reduce' per-op tolerances of DESIGN.md section 4; 'close' rtol/atol 1e-5 (fp32 results of
reductions / gradients that accumulate); 'bf16' = bf16 contract (reference evaluated in fp32 on the same bf16 values,
rounded once, rtol 1e-2).
"""

F = dict(lo=1, hi=7, feat=(3,), dtype='f32')

CASES = [
    # a13-a16 conversions
    ('conv_ties_f32', 'conversions', dict(seed=1, B=9, **F, distinct=False, fills=(0, 2.5)), 'exact', None),
    ('conv_distinct_bf16', 'conversions', dict(seed=2, B=8, lo=1, hi=0, feat=(5,), dtype='bf16', distinct=True, fills=(0, -1)), 'exact', None),
    ('conv_featureless_i64', 'conversions', dict(seed=3, B=6, lo=2, hi=0, feat=(), dtype='i64', distinct=True), 'exact', None),
    ('conv_wide_f16', 'conversions', dict(seed=4, B=5, lo=3, hi=0, feat=(2, 40), dtype='f16', distinct=True), 'exact', None),
    # a24-a28 selects
    ('selects_distinct_f32', 'selects', dict(seed=5, B=7, lo=2, hi=0, feat=(3,), dtype='f32', distinct=True), 'exact', None),
    ('selects_ties_i32', 'selects', dict(seed=6, B=10, lo=2, hi=5, feat=(), dtype='i32', distinct=False), 'exact', None),
    ('selects_wide_bf16', 'selects', dict(seed=7, B=4, lo=3, hi=0, feat=(72,), dtype='bf16', distinct=True, shifts=(2, -7)), 'exact', None),
    # a2-a12, a18 helpers, views, masks
    ('meta_distinct_f32', 'metadata', dict(seed=8, B=7, lo=1, hi=0, feat=(3,), dtype='f32', distinct=True), 'exact', None),
    ('meta_ties_f32', 'metadata', dict(seed=9, B=12, lo=1, hi=5, feat=(2,), dtype='f32', distinct=False), 'exact', None),
    ('meta_featureless_i64', 'metadata', dict(seed=10, B=6, lo=2, hi=0, feat=(), dtype='i64', distinct=True), 'exact', None),
    ('meta_single', 'metadata', dict(seed=11, B=1, lo=4, hi=0, feat=(2,), dtype='f32', distinct=True), 'exact', None),
    # a17 __getitem__ / __setitem__
    ('getitem_f32', 'getitem', dict(seed=12, B=7, lo=2, hi=0, feat=(3,), dtype='f32'), 'exact', None),
    ('getitem_wide_storage', 'getitem', dict(seed=13, B=6, lo=1, hi=0, feat=(2,), dtype='f32', extra_width=3), 'exact', None),
    ('getitem_featureless_i64', 'getitem', dict(seed=14, B=5, lo=2, hi=0, feat=(), dtype='i64'), 'exact', None),
    ('getitem_rows_160B', 'getitem', dict(seed=15, B=5, lo=2, hi=0, feat=(40,), dtype='f32'), 'exact', None),
    ('setitem_f32', 'setitem', dict(seed=16, B=7, lo=2, hi=0, feat=(3,), dtype='f32'), 'exact', None),
    ('setitem_wide_storage', 'setitem', dict(seed=17, B=6, lo=1, hi=0, feat=(2,), dtype='f32', extra_width=2), 'exact', None),
    ('setitem_featureless_i64', 'setitem', dict(seed=18, B=5, lo=2, hi=0, feat=(), dtype='i64'), 'exact', None),
    ('setitem_rows_256B', 'setitem', dict(seed=19, B=4, lo=3, hi=0, feat=(128,), dtype='bf16'), 'exact', None),
    # a19-a23 reductions
    ('reduce_f32', 'reductions', dict(seed=20, S=11, lo=1, hi=9, feat=(4,), dtype='f32', grad=True), 'reduce', None),
    ('reduce_empty_segments', 'reductions', dict(seed=21, S=13, lo=0, hi=5, feat=(3,), dtype='f32', grad=True), 'reduce', None),
    ('reduce_flat_f32', 'reductions', dict(seed=22, S=40, lo=1, hi=30, feat=(), dtype='f32', grad=True), 'reduce', None),
    ('reduce_f64', 'reductions', dict(seed=23, S=6, lo=1, hi=50, feat=(2,), dtype='f64'), 'reduce', None),
    ('reduce_long_wide', 'reductions', dict(seed=24, S=5, lo=100, hi=700, feat=(70,), dtype='f32'), 'reduce', None),
    ('reduce_bf16_contract', 'reductions', dict(seed=25, S=9, lo=1, hi=300, feat=(16,), dtype='bf16',
                                                ops=('sum', 'mean', 'max', 'min', 'logsumexp')), 'bf16', dict(upcast=True)),
    ('seg_f32', 'seg', dict(seed=26, B=5, lo=2, hi=0, feat=(3,), dtype='f32'), 'close', None),
    # gradients of conversions / selects / getitem
    ('gradients', 'gradients', dict(seed=27, B=6, lo=3, hi=0, feat=(2,)), 'close', None),
    # (f) rows
    ('constructors_f32', 'constructors', dict(seed=28, B=6, lo=1, hi=0, feat=(3,), dtype='f32'), 'exact', None),
    ('compose_f32', 'compose', dict(seed=29, B=4, lo=1, hi=0, feat=(2,), dtype='f32'), 'exact', None),
    ('scatter_f32', 'scatter', dict(seed=30, M=9, K=40, feat=(3,), dtype='f32'), 'reduce', None),
]

# larger shapes: live reference only (run on the GPU box against the reference's CUDA path; nothing frozen)
LIVE_CASES = [
    ('live_conv_bf16_H256', 'conversions', dict(seed=40, B=300, lo=1, hi=90, feat=(256,), dtype='bf16', distinct=False, fills=(0,)), 'exact', None),
    ('live_conv_distinct_H64', 'conversions', dict(seed=41, B=257, lo=1, hi=0, feat=(64,), dtype='f32', distinct=True, fills=(1.5,)), 'exact', None),
    ('live_conv_ids', 'conversions', dict(seed=42, B=3000, lo=1, hi=64, feat=(), dtype='i64', distinct=False), 'exact', None),
    ('live_selects_H128', 'selects', dict(seed=43, B=200, lo=2, hi=70, feat=(128,), dtype='bf16', distinct=False, shifts=(1, -3, 17)), 'exact', None),
    ('live_selects_distinct_ids', 'selects', dict(seed=44, B=150, lo=2, hi=0, feat=(), dtype='i64', distinct=True), 'exact', None),
    ('live_meta_distinct', 'metadata', dict(seed=45, B=500, lo=1, hi=0, feat=(4,), dtype='f32', distinct=True), 'exact', None),
    ('live_meta_ties', 'metadata', dict(seed=46, B=9000, lo=1, hi=64, feat=(), dtype='i64', distinct=False), 'exact', None),
    ('live_getitem_H512', 'getitem', dict(seed=47, B=120, lo=1, hi=0, feat=(512,), dtype='bf16'), 'exact', None),
    ('live_getitem_wide_storage_ids', 'getitem', dict(seed=48, B=400, lo=1, hi=0, feat=(), dtype='i64', extra_width=5), 'exact', None),
    ('live_setitem_H512', 'setitem', dict(seed=49, B=120, lo=1, hi=0, feat=(512,), dtype='bf16'), 'exact', None),
    ('live_setitem_wide_storage_f32', 'setitem', dict(seed=50, B=300, lo=1, hi=0, feat=(3,), dtype='f32', extra_width=4), 'exact', None),
    ('live_reduce_f32_H256', 'reductions', dict(seed=51, S=700, lo=0, hi=200, feat=(256,), dtype='f32', grad=True), 'reduce', None),
    # >= 32768 segments of per-token scalars: the warp-per-32-segments kernels, forward and backward (with empty segments)
    ('live_reduce_flat', 'reductions', dict(seed=52, S=40000, lo=0, hi=64, feat=(), dtype='f32', grad=True), 'reduce', None),
    ('live_reduce_flat_long_tail', 'reductions', dict(seed=59, S=33000, lo=1, hi=9, feat=(), dtype='f32', grad=True, scale=0.5), 'reduce', None),
    ('live_reduce_flat_f64', 'reductions', dict(seed=60, S=34000, lo=0, hi=20, feat=(), dtype='f64', grad=True), 'reduce', None),
    ('live_reduce_bf16_H1024', 'reductions', dict(seed=53, S=300, lo=1, hi=512, feat=(1024,), dtype='bf16',
                                                  ops=('sum', 'mean', 'max', 'min', 'logsumexp')), 'bf16', dict(upcast=True)),
    ('live_seg', 'seg', dict(seed=54, B=40, lo=2, hi=0, feat=(16,), dtype='f32'), 'close', None),
    ('live_gradients', 'gradients', dict(seed=55, B=60, lo=3, hi=0, feat=(32,)), 'close', None),
    ('live_constructors', 'constructors', dict(seed=56, B=50, lo=1, hi=0, feat=(8,), dtype='f32'), 'exact', None),
    ('live_compose', 'compose', dict(seed=57, B=30, lo=1, hi=0, feat=(8,), dtype='f32'), 'exact', None),
    ('live_scatter', 'scatter', dict(seed=58, M=200, K=5000, feat=(16,), dtype='f32'), 'reduce', None),
]


# a seeded sweep over shapes / dtypes / row widths (vector widths 1 B .. 32 B, narrow and wide kernels) against the live reference
def _sweep():
    import random
    rnd = random.Random(2024)
    feats = [(), (1,), (3,), (4,), (6,), (16,), (31,), (32,), (33,), (64,), (96,), (130,), (256,), (2, 24)]
    dtypes = ['f32', 'bf16', 'f16', 'i64', 'i32', 'f64']
    out = []
    for k in range(18):
        feat, dt = rnd.choice(feats), rnd.choice(dtypes)
        b = rnd.choice([1, 2, 5, 17, 33, 64, 150, 700])
        distinct = b <= 150 and rnd.random() < 0.6
        hi = rnd.choice([3, 9, 40, 130])
        base = dict(seed=1000 + k, B=b, lo=1, hi=0 if distinct else hi, feat=feat, dtype=dt)
        fn = ['conversions', 'selects', 'getitem', 'setitem', 'metadata'][k % 5]
        kw = dict(base)
        if fn in ('conversions', 'selects', 'metadata'):
            kw['distinct'] = distinct
            if fn == 'selects':
                kw['lo'] = 2
        else:
            kw['hi'] = 0                    # getitem / setitem scenarios draw distinct lengths themselves
            kw['B'] = min(b, 150)
            kw['extra_width'] = rnd.choice([0, 0, 2])
        out.append((f'sweep{k:02d}_{fn}_{dt}_{"x".join(map(str, feat)) or "scalar"}_B{kw["B"]}', fn, kw, 'exact', None))
    return out


LIVE_CASES += _sweep()
