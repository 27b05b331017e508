"""Device-side code sampling (SNGan.sample_codes -> tf.random_normal in the reference's graph, DeepLearning/my_sngan.py:122-124).
CPU: the numpy restatement of Philox-4x32-10 against the algorithm's published known-answer vectors (Random123 kat_vectors).
GPU: the CUDA kernel against that restatement -- Philox words bit for bit, normals to fp32 rounding, moments / tails of the
distribution -- and the engine drawing its own codes inside the captured step."""
import numpy as np
import pytest
import torch

from oracle import philox as ph

KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def test_philox_oracle_known_answer_vectors():
    for ctr, key, want in KAT:
        got = ph.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(x) for x in got) == want
    z = ph.sample_normal(200000, 1234, draw=7)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert not np.array_equal(ph.sample_words(64, 5, 0), ph.sample_words(64, 5, 1))      # the draw counter moves the stream


@pytest.mark.gpu
def test_sample_normal_kernel_matches_oracle(cuda):
    from mmdgan_b200 import kernels as K
    for n, seed, draw in ((256 * 128, 0x1234_5678_9ABC_DEF0, 0), (1003, 7, 5), (4, 2 ** 63 + 11, 2 ** 40 + 3)):
        out = torch.empty(n, device=cuda)
        raw = torch.zeros(n, dtype=torch.int32, device=cuda)
        ctr = torch.tensor([draw], dtype=torch.int64, device=cuda)
        K.sample_normal(out, seed, ctr, raw=raw)
        words = raw.cpu().numpy().view(np.uint32)
        assert np.array_equal(words, ph.sample_words(n, seed, draw))                      # bit exact
        ref = ph.sample_normal(n, seed, draw)
        assert np.abs(out.cpu().numpy().astype(np.float64) - ref).max() < 2e-5            # logf / sincospif: fp32 rounding
    big = torch.empty(1 << 22, device=cuda)
    K.sample_normal(big, 99, torch.tensor([1], dtype=torch.int64, device=cuda))
    z = big.double()
    assert abs(float(z.mean())) < 2e-3 and abs(float(z.var()) - 1.0) < 3e-3
    assert abs(float((z ** 3).mean())) < 1e-2 and abs(float((z ** 4).mean()) - 3.0) < 3e-2
    assert abs(float((z.abs() > 3.0).double().mean()) - 0.0026998) < 2e-4                   # tail mass beyond 3 sigma
    assert torch.isfinite(big).all()


@pytest.mark.gpu
def test_engine_draws_its_codes_on_the_device(cuda):
    """step(data) with no codes: z ~ N(0, 1) is drawn inside the step (one Philox draw per step; the captured graph produces
    fresh codes on every replay), and equals what SNGan.sample_codes hands out for the same torch seed -- so feeding those
    codes back explicitly reproduces the step bit for bit."""
    from oracle import architectures as oa
    from oracle import net as onet
    from mmdgan_b200.engine import SNGanEngine
    from mmdgan_b200.DeepLearning.my_sngan import SNGan
    arch = oa.tiny(act_k=2.6)
    B = 16
    torch.manual_seed(123)
    eng_dev = SNGanEngine(arch, B, loss_type='rep', use_graph=True, seed=3)
    eng_host = SNGanEngine(arch, B, loss_type='rep', use_graph=True, seed=3)
    mdl = SNGan(arch, num_class=0, loss_type='rep', optimizer='adam')
    seen = []
    for it in range(4):
        data, _ = onet.synthetic_batch(arch, B, seed=40 + it)
        l_dev = eng_dev.step(data)                                  # codes drawn on the device
        codes = mdl.sample_codes(B)['x']                            # the same generator, host side of the API
        seen.append(codes.clone())
        assert torch.equal(eng_dev._dev_code.cpu(), codes)
        l_host = eng_host.step(data, codes)
        assert l_dev == l_host, (it, l_dev, l_host)
    assert eng_dev._graphs is not None
    assert all(not torch.equal(seen[0], s) for s in seen[1:])       # a new draw every step, also from the replayed graph
    z = torch.cat(seen).double()
    assert abs(float(z.mean())) < 0.05 and abs(float(z.std()) - 1.0) < 0.05
