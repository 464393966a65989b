"""Worker of tests/test_gpu_multirank.py: one process per GPU under torchrun.  Every rank holds a full replica, takes its
rows of the global batch, runs optimizer steps whose gradients meet in ONE parrot_comm_allreduce
(NCCL through the C ABI), and rank 0 writes parameters / global costs for the comparison with a single process."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CFG = dict(input_dim=24, output_dim=63, rnn_h_dim=128, readouts_dim=128, num_characters=43, attention_size=10,
           encoder_dim=64, encoder_type='bidirectional', weak_feedback=True, attention_alignment=0.4)
B_GLOBAL, T, U, STEPS = 16, 24, 16, 2


def make_model(dev):
    from parrot_b200 import Parrot
    m = Parrot(device=dev, encoder_time_axis=1, **CFG)     # literal axis 0 mixes batch rows
    m.initialize(seed=3, gain=0.5)
    return m


def batches():
    from parrot_b200.synthetic import make_batch
    return [make_batch(CFG, B_GLOBAL, T, U, seed=50 + i) for i in range(STEPS)]


def run(model, shard, world, rank):
    from parrot_b200 import parallel
    from parrot_b200.algorithms import Adam, CompositeRule, GradientDescent, StepClipping
    algo = GradientDescent(model=model, step_rule=CompositeRule([StepClipping(0.5), Adam(1e-3)]))
    costs = []
    for bt in batches():
        b = parallel.shard_batch(bt, rank, world) if shard else bt
        batch = dict(features=b['features'], features_mask=b['features_mask'], labels=b['labels'],
                     labels_mask=b['labels_mask'], start_flag=1.0)
        algo.process_batch(batch, b['features'].shape[1])
        costs.append(algo.global_cost())
    torch.cuda.synchronize()
    return np.array(costs), model.flat_params.detach().cpu().numpy().copy(), algo.stats.cpu().numpy().copy()


def main():
    from parrot_b200 import parallel
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    model = make_model(dev)
    costs, params, stats = run(model, True, world, rank)
    comm = parallel.get_comm()
    info = comm.info() if comm is not None else (1, 0, 0)
    if rank == 0:
        np.savez(sys.argv[1], costs=costs, params=params, stats=stats, info=np.array(info))
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
