"""CPU model of the 3xTF32 GEMMs of the wide flavour: tf32 hi/lo split, K = 8 products per MMA summed exactly, then
added into an FP32 accumulator with round-toward-zero (RZ) or round-to-nearest (RN), in the MMA orders the kernels use
("interleaved": per sub-step lo*hi, hi*lo, hi*hi) and two alternatives.  Prints max-norm error, relative rms error and
the mean signed relative error (the systematic shrink that SweepArgs::comp_* corrects) against float64.
"""
import numpy as np
rng = np.random.RandomState(0)
def tf32_rn(x):
    b = x.astype(np.float32).view(np.uint32)
    b = ((b + np.uint32(0x1000)) & np.uint32(0xffffe000))
    return b.view(np.float32)
def tf32_trunc(x):
    b = x.astype(np.float32).view(np.uint32) & np.uint32(0xffffe000)
    return b.view(np.float32)
def rz32(x64):
    # round float64 toward zero to float32
    f = x64.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(x64)
    f2 = np.nextafter(f, np.float32(0))
    return np.where(over, f2, f)
def rn32(x64): return x64.astype(np.float32)

def gemv(A, W, order, lo_mode, acc_round, nsub, two_acc=False):
    # A [R,K], W [N,K] fp32. returns [R,N]
    R_, K = A.shape; N = W.shape[0]
    Ah = tf32_rn(A); Al = (A - Ah).astype(np.float32)
    Wh = tf32_rn(W); Wl = (W - Wh).astype(np.float32)
    if lo_mode == 'trunc': Al = tf32_trunc(Al); Wl = tf32_trunc(Wl)
    else: Al = tf32_rn(Al); Wl = tf32_rn(Wl)
    ks = K // 8
    # k-step order per substep: groups g=0..3, substep j: kstep = g*nsub + j
    seq = []
    for j in range(nsub):
        kk = [g * nsub + j for g in range(4)]
        seq.append(kk)
    acc = np.zeros((R_, N), np.float32); acc2 = np.zeros((R_, N), np.float32)
    def mma(acc, X, Y, k):
        p = X[:, 8*k:8*k+8].astype(np.float64) @ Y[:, 8*k:8*k+8].astype(np.float64).T
        return acc_round(acc.astype(np.float64) + p)
    if order == 'interleaved':
        for kk in seq:
            for k in kk: acc = mma(acc, Al, Wh, k)
            for k in kk: acc = mma(acc, Ah, Wl, k)
            for k in kk: acc = mma(acc, Ah, Wh, k)
    elif order == 'corr_first':
        for kk in seq:
            for k in kk: acc = mma(acc, Al, Wh, k)
            for k in kk: acc = mma(acc, Ah, Wl, k)
        for kk in seq:
            for k in kk: acc = mma(acc, Ah, Wh, k)
    elif order == 'two_acc':
        for kk in seq:
            for k in kk: acc2 = mma(acc2, Al, Wh, k)
            for k in kk: acc2 = mma(acc2, Ah, Wl, k)
            for k in kk: acc = mma(acc, Ah, Wh, k)
        acc = (acc + acc2).astype(np.float32)
    return acc

for K in (64, 128):
    R_ = 512; N = K
    A = rng.randn(R_, K).astype(np.float32)
    W = (rng.rand(N, K).astype(np.float32) * 2 - 1) / np.sqrt(K)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    fp32 = np.zeros((R_, N), np.float32)
    for k in range(K): fp32 = (fp32 + A[:, k:k+1] * W[None, :, k]).astype(np.float32)
    scale = np.abs(ref).max()
    def err(x): 
        d = x.astype(np.float64) - ref
        return np.abs(d).max()/scale, np.sqrt((d**2).mean())/np.sqrt((ref**2).mean()), (d*np.sign(ref)).mean()/np.abs(ref).mean()
    print('K', K, 'fp32 sequential FMA-ish', ['%.2e' % v for v in err(fp32)])
    for order in ('interleaved','corr_first','two_acc'):
        for lo_mode in ('trunc','rn'):
            for name, rnd in (('RZ', rz32), ('RN', rn32)):
                o = gemv(A, W, order, lo_mode, rnd, K // 32)
                print('  ', order, lo_mode, name, ['%.2e' % v for v in err(o)])
