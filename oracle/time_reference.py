"""TEST INFRASTRUCTURE ONLY (build container; /root/reference is not on the GPU box).  Times the UNMODIFIED reference's refine
iteration (decode_embedding + AvoidCollLoss + backward + Adam.step, refine_traffic_optim.py:185-218, weights requires_grad as the
drivers leave them) next to the oracle port on the SAME sample of the bench workload (1 scene x 32 agents x 20 steps), same
thread count -- the port-vs-reference factor quoted in BASELINE.md.    python oracle/time_reference.py"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims, gen_golden as GG, strive_oracle as O      # noqa: E402
from strive_b200 import synth                                            # noqa: E402
import bench                                                              # noqa: E402


def main():
    threads = int(os.environ.get('THREADS', os.cpu_count() or 1))
    torch.set_num_threads(threads)
    ref_shims.install()
    raster, dx, sd, sc = bench.make_workload(0, scenes=1)
    env = ref_shims.make_map_env(raster, dx)
    model = GG.quiet(ref_shims.make_ref_model, nfuture=20)
    model.load_state_dict(sd, strict=False)
    model.train()
    from losses.adv_gen_nusc import AvoidCollLoss
    FT, iters = bench.WORK['FT'], int(os.environ.get('ITERS', '2'))
    NA = sc['z'].size(0)

    def run_ref(n_iter):
        z = sc['z'].clone().requires_grad_(True)
        opt = torch.optim.Adam([z], lr=bench.LR)
        lw_un = model.get_att_normalizer().unnormalize(sc['lw'])
        lf = GG.quiet(AvoidCollLoss, bench.REFINE_W, lw_un, sc['map_idx'][sc['batch']], env, z.clone().detach(), veh_coll_buffer=0.2)
        t0 = time.perf_counter()
        for _ in range(n_iter):
            opt.zero_grad()
            model.zero_grad()
            fut = GG.decode_ref(model, env, sc, z, FT)
            ld = lf(model.get_normalizer().unnormalize(fut), z, (sc['prior_mu'], sc['prior_var']))
            ld['loss'].backward()
            opt.step()
        return time.perf_counter() - t0

    def run_port(n_iter):
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        z = sc['z'].clone().requires_grad_(True)
        opt = torch.optim.Adam([z], lr=bench.LR)
        lw_un = O.unnorm_att(sc['lw'])
        t0 = time.perf_counter()
        for _ in range(n_iter):
            opt.zero_grad()
            for p in sdg.values():
                p.grad = None
            fut = O.decode(sdg, z, sc['map_feat'], sc['past_feat'], sc['past'][:, -1, :], sc['lw'], sc['sem'], sc['ptr'], sc['edge_index'], sc['map_idx'],
                           raster, dx, FT)
            ld = O.avoid_coll_loss(O.unnorm_state(fut), z, (sc['prior_mu'], sc['prior_var']), sc['z'], bench.REFINE_W, lw_un, sc['map_idx'][sc['batch']],
                                   None, raster, dx, veh_coll_buffer=0.2)
            ld['loss'].backward()
            opt.step()
        return time.perf_counter() - t0

    run_ref(1); run_port(1)
    tr, tp = run_ref(iters), run_port(iters)
    u = NA * FT * iters
    print('unmodified reference: %.1f units/s (%.2f s/iter) | oracle port: %.1f units/s (%.2f s/iter) | port/reference = %.2fx | %d agents x %d steps, %d torch threads'
          % (u / tr, tr / iters, u / tp, tp / iters, tr / tp, NA, FT, threads))


if __name__ == '__main__':
    main()
