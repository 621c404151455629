"""Small end-to-end run for compute-sanitizer (memcheck): every kernel of the path at tiny sizes."""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import myzkp_b200 as mz
import myzkp_oracle as o

ctx = mz.Context(0)
alpha, u = 123456789, 5
for n, c, seg, baa in ((300, 0, 0, -1), (1500, 12, 3, -1), (5000, 16, 0, 2), (70000, 20, 0, -1)):
    ctx.srs_generate(alpha, n)
    coefs = [(i * 7919 + 13) % o.R_MOD for i in range(n)]
    ctx.set_msm_params(c, seg)
    ctx.set_baa_rounds(baa)
    assert ctx.commit(coefs) == o.expected_commit(coefs, alpha)
    assert ctx.open(coefs, u) == o.expected_open(coefs, u, alpha)
ctx.set_msm_params(0, 0)
ctx.set_baa_rounds(-1)
ctx.srs_generate(alpha, 16)
print(ctx.gemini_fold_commit(list(range(1, 17)), [2, 3, 4, 5])[:1])
print(ctx.batch_open(list(range(1, 17)), [7, 8, 9])[0])
print(ctx.prove_degree_bound(list(range(1, 9)), 8) is not None)
print("sanitize run ok")
