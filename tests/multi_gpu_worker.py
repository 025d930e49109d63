"""Worker for tests/test_gpu_multi.py: run under torchrun (one rank per GPU).  Each rank runs its contiguous
block of chains + its block of a parameter sweep; rank 0 saves the gathered results for comparison with a
single-GPU run of the very same chains."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odelib_b200.rhat import shard_bounds, sharded_mcmc  # noqa: E402
from tests.helpers import device_model, golden, prior_draws  # noqa: E402


def main(out_path):
    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dm, _ = device_model("two_i", device=local)
    g = golden("two_i")
    C, nits = 96, 120
    rng = np.random.default_rng(42)
    theta0 = g["chain_def_s0_theta0"] * np.exp(0.05 * rng.standard_normal((C, 5)))
    res, rh, (lo, hi) = sharded_mcmc(dm, torch.from_numpy(theta0).cuda(), nits=nits, seed=9, rng_mode="philox")
    samples = res["samples"]
    theta = prior_draws("two_i", 5000, seed=4)
    slo, shi = shard_bounds(len(theta), ws, rank)
    sw = dm.sweep(torch.from_numpy(theta[slo:shi]).cuda(), solver="auto")
    chi = sw["chi"]
    if ws > 1:
        # gather the per-rank blocks only to let the test compare them; the data path itself has no collective
        parts = [None] * ws
        dist.all_gather_object(parts, (lo, samples.cpu().numpy(), slo, chi.cpu().numpy()))
    else:
        parts = [(lo, samples.cpu().numpy(), slo, chi.cpu().numpy())]
    # the reference's own surface, sharded by the facade itself (ModelFramework(distributed=True)): every rank makes the
    # same calls and gets the complete frames
    from tests.test_facade_host import make_model
    m = make_model("two_i", distributed=True, device=local)
    np.random.seed(11)
    sv = m.fit_survey(samples=3001)
    np.random.seed(12)
    post = m.MCMC(chain_inits=7, iterations_per_chain=80, fitsurvey_samples=3001, sd_fitdistance=6.0, print_report=False)
    rh_f, ess_f = np.array(list(m.rhat.values())), np.array(list(m.ess.values()))
    summ = m.MCMC(chain_inits=[m.get_parameters(as_dict=True)] * 5, iterations_per_chain=60, print_report=False,
                  posterior="summary", rng="philox")
    one = m.MCMC(chain_inits=[{}], iterations_per_chain=40, print_report=False)      # fewer chains than ranks
    # a NON-distributed facade on a handle that has joined the communicator (what bench.py does: its MCMC leg joins the
    # model's handle, then rank 0 alone runs ModelFramework.MCMC): rank 0's R-hat must not start a collective
    m1 = make_model("two_i", device=local)
    m1._device().comm_init()
    if rank == 0:
        solo = m1.MCMC(chain_inits=[m1.get_parameters(as_dict=True)] * 6, iterations_per_chain=40, print_report=False)
        assert len(solo) == 6 * 19 and np.all(np.isfinite(list(m1.rhat.values())))
    if ws > 1:
        frames = [None] * ws
        dist.all_gather_object(frames, (sv.to_numpy(), post.to_numpy()))
        for f in frames[1:]:                                      # every rank holds the same complete result
            np.testing.assert_array_equal(f[0], frames[0][0])
            np.testing.assert_array_equal(f[1], frames[0][1])
    if rank == 0:
        parts.sort(key=lambda p: p[0])
        np.savez(out_path, samples=np.concatenate([p[1] for p in parts]), chi=np.concatenate([p[3] for p in parts]),
                 rhat=rh, world=ws, facade_survey=sv.to_numpy(), facade_post=post.to_numpy(), facade_rhat=rh_f,
                 facade_ess=ess_f, facade_best=np.array(list(summ.best.values()) + [summ.best_chi, summ.acceptance_ratio]),
                 facade_one=one.to_numpy())
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
