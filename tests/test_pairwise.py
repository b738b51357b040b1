"""Pairwise additive decoder: oracle vs the reference's own output (CPU), and the CUDA
kernel through the C ABI vs the oracle (GPU) -- bit-exact, the additions are done in the reference's order."""
import json
import os

import numpy as np
import pytest

from oracle import qinco_oracle as orc
from qinco_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pairwise_ivf.npz")


def load_case():
    z = np.load(GOLDEN)
    kw = json.loads(str(z["cfg"]))
    book, comb, imap = synth.make_pairwise_tables(seed=int(z["seed"]), **kw)
    return kw, book, comb, imap, z


def test_oracle_matches_reference_output():
    """tests/golden/pairwise_ivf.npz holds PairwiseDecoderIVF.forward of the UNMODIFIED reference (oracle/make_golden.py)."""
    kw, book, comb, imap, z = load_case()
    out = orc.pairwise_decode(book, comb, imap, kw["K"], z["codes"], z["ivf_codes"])
    assert out.dtype == np.float32 and np.array_equal(out, z["out_ref"])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from qinco_b200 import _lib
    from qinco_b200.pairwise import PairwiseDecoderIVF
    kw, book, comb, imap, _ = load_case()
    sd = dict(codebook_MKD=book, combine_mvals_m=comb, ivf_code_map=imap)
    with pytest.raises(RuntimeError):
        PairwiseDecoderIVF(sd, K=kw["K"], M=kw["M"], device="cpu")
    with pytest.raises(_lib.QbError):
        PairwiseDecoderIVF(sd, K=kw["K"], M=kw["M"], device="cuda:0")


@pytest.mark.gpu
def test_kernel_bit_exact_vs_reference_and_oracle():
    import torch
    from qinco_b200.pairwise import PairwiseDecoderIVF
    kw, book, comb, imap, z = load_case()
    dec = PairwiseDecoderIVF(dict(codebook_MKD=torch.from_numpy(book), combine_mvals_m=torch.from_numpy(comb),
                                  ivf_code_map=torch.from_numpy(imap)), K=kw["K"], M=kw["M"])
    try:
        out = dec(torch.from_numpy(z["codes"]), torch.from_numpy(z["ivf_codes"]))
        torch.cuda.synchronize()
        assert out.dtype == torch.float32 and tuple(out.shape) == z["out_ref"].shape
        assert np.array_equal(out.cpu().numpy(), z["out_ref"])
        # empty and single-vector batches, int32 inputs (search_tasks.py:428-445 passes int32 codes)
        assert dec(torch.zeros((kw["M"], 0), dtype=torch.int64), torch.zeros(0, dtype=torch.int64)).shape == (0, kw["D"])
        one = dec(torch.from_numpy(z["codes"][:, :1]).int(), torch.from_numpy(z["ivf_codes"][:1]).int())
        assert np.array_equal(one.cpu().numpy(), z["out_ref"][:1])
        with pytest.raises(IndexError):
            dec(torch.full((kw["M"], 2), kw["K"], dtype=torch.int64), torch.zeros(2, dtype=torch.int64))
        with pytest.raises(IndexError):
            dec(torch.zeros((kw["M"], 2), dtype=torch.int64), torch.full((2,), kw["ivf_K"], dtype=torch.int64))
        assert dec.launch_count >= 2
    finally:
        dec.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [dict(D=768, M=8, K=16, Mt=16, ivf_K=64), dict(D=96, M=16, K=8, Mt=32, ivf_K=16),
                                   dict(D=4, M=1, K=3, Mt=1, ivf_K=1)])
def test_kernel_shapes_vs_oracle(shape):
    """Contriever / Deep1B-like dims (small K so the tables stay small), > 16 tables (two load batches), tiny dims; ragged n."""
    import torch
    from qinco_b200.pairwise import PairwiseDecoderIVF
    book, comb, imap = synth.make_pairwise_tables(seed=9, **shape)
    rng = np.random.default_rng(3)
    n = 1237
    codes = rng.integers(0, shape["K"], (shape["M"], n)).astype(np.int64)
    ivf = rng.integers(0, shape["ivf_K"], n).astype(np.int64)
    ref = orc.pairwise_decode(book, comb, imap, shape["K"], codes, ivf)
    dec = PairwiseDecoderIVF(dict(codebook_MKD=book, combine_mvals_m=comb, ivf_code_map=imap), K=shape["K"], M=shape["M"])
    try:
        out = dec(torch.from_numpy(codes), torch.from_numpy(ivf))
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ref)
    finally:
        dec.close()
