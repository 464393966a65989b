"""Training entry point (reference train.py:1-200), same flags (parrot_b200/utils.py), same call order:
parse -> streams -> Parrot(**parrot_args).initialize() -> compute_cost -> GradientDescent(
CompositeRule([StepClipping(10 * grad_clip), Adam(lr)])) -> main loop with monitoring / checkpoints.

The Blocks MainLoop machinery is replaced by a plain loop; data-parallel runs launch one process per
GPU with torchrun (rank/world from the environment) and add one NCCL allreduce per step.

    RESULTS_DIR=/tmp/res python train.py --batch_size 8 --seq_size 50 --rnn_h_dim 64 --steps 20
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 train.py --batch_size 512 ...
"""
import os
import pickle
import time

import numpy
import torch

from parrot_b200 import Parrot, parallel
from parrot_b200.algorithms import Adam, CompositeRule, GradientDescent, StepClipping
from parrot_b200.datasets import SyntheticVoice, parrot_stream
from parrot_b200.utils import train_parse


class LearningRateSchedule(object):
    """extensions.py:83-150 of the reference, minus the Blocks plumbing: halve the learning rate when the tracked
    validation cost has not improved for ``patience`` checks (or is NaN), reload the best parameters, zero the Adam
    buffers, and ask for the end of training after ``num_cuts`` cuts (train.py:175-181: patience 10, num_cuts 5)."""

    def __init__(self, adam, algorithm, model, path, patience=5, num_cuts=3, cut_size=.5):
        self.adam, self.algorithm, self.model, self.path = adam, algorithm, model, path
        self.patience, self.num_cuts, self.cut_size = patience, num_cuts, cut_size
        self.counter = self.count_cuts = 0
        self.best_value = numpy.inf

    def do(self, current_value, rank=0, world=1):
        """Returns True when training should finish.  Every rank calls this with the same (global) value."""
        if current_value is None:
            return False
        if current_value < self.best_value:
            self.best_value, self.counter = current_value, 0
        else:
            self.counter += 1
        if numpy.isnan(current_value):
            self.counter = self.patience + 1
        if self.counter < self.patience:
            return False
        self.counter = 0
        self.count_cuts += 1
        if rank == 0 and os.path.exists(self.path):
            self.model.set_parameter_values(dict(numpy.load(self.path)))
        if world > 1:                                  # replicas stay identical: rank 0's reloaded parameters
            torch.distributed.broadcast(self.model.flat_params, 0)
            self.model.mark_dirty()
        self.algorithm.zero_buffers()
        self.adam.learning_rate = float(self.cut_size * self.adam.learning_rate)
        return self.count_cuts >= self.num_cuts


def main(argv=None):
    args = train_parse(argv)
    rank, world, local = parallel.init_from_env()
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    exp_name, save_dir = args.experiment_name, args.save_dir
    if rank == 0:
        for sub in ('config', 'pkl', 'progress', 'samples'):
            os.makedirs(os.path.join(save_dir, sub), exist_ok=True)
        print('Saving config ...')
        with open(os.path.join(save_dir, 'config', exp_name + '.pkl'), 'wb') as f:   # train.py:25-28
            pickle.dump(args, f)
        print('Finished saving.')

    assert args.batch_size % world == 0, 'global batch must divide by the number of ranks'
    labels_type = args.labels_type
    if labels_type not in ('text', 'unaligned_phonemes'):
        # frame-aligned label types bypass the attention path this package implements (model.py:577-583)
        if rank == 0:
            print("labels_type %r is not on the attention path; using 'text' (sequence-level labels)" % labels_type)
        labels_type = 'text'
    dataset = SyntheticVoice(num_examples=max(256, 4 * args.batch_size), output_dim=args.output_dim,
                             num_characters=args.num_characters, num_speakers=args.num_speakers,
                             seed=args.seed)
    train_stream = parrot_stream(args.dataset, args.use_speaker, ('train',), args.batch_size,
                                 noise_level=args.feedback_noise_level, labels_type=labels_type,
                                 seq_size=args.seq_size, dataset=dataset, seed=args.seed, epochs=10 ** 6)
    # held-out utterances for the validation monitor (another seed of the synthetic source: disjoint from training)
    valid_dataset = SyntheticVoice(num_examples=max(64, 4 * args.batch_size), output_dim=args.output_dim,
                                   num_characters=args.num_characters, num_speakers=args.num_speakers,
                                   seed=args.seed + 7919)
    valid_stream_args = dict(noise_level=None if args.feedback_noise_level is None else 0.,
                             labels_type=labels_type, seq_size=args.seq_size, dataset=valid_dataset,
                             seed=args.seed + 1)

    parrot_args = {                                                            # train.py:57-77
        'input_dim': args.input_dim, 'output_dim': args.output_dim, 'rnn_h_dim': args.rnn_h_dim,
        'readouts_dim': args.readouts_dim, 'weak_feedback': args.weak_feedback,
        'full_feedback': args.full_feedback, 'feedback_noise_level': args.feedback_noise_level,
        'layer_norm': args.layer_norm, 'use_speaker': args.use_speaker,
        'num_speakers': args.num_speakers, 'speaker_dim': args.speaker_dim,
        'which_cost': args.which_cost, 'num_characters': args.num_characters,
        'attention_type': args.attention_type, 'attention_alignment': args.attention_alignment,
        'encoder_type': args.encoder_type, 'raw_output': args.raw_output, 'name': 'parrot'}
    # data-parallel parity needs the encoder to recur over text positions, not over the batch axis
    # (SURVEY D4 / 8e): literal mode mixes batch rows inside the encoder.
    parrot = Parrot(encoder_time_axis=0 if world == 1 else 1, **parrot_args)
    parrot.initialize(seed=args.seed)                                          # train.py:79-80 (N(0, 0.01), 0)

    step_rule = CompositeRule([StepClipping(10. * args.grad_clip), Adam(args.learning_rate)])   # train.py:100-101
    algorithm = GradientDescent(cost=None, parameters=parrot.parameters, step_rule=step_rule, model=parrot)

    if args.load_experiment:                                                   # train.py:136-139
        path = os.path.join(save_dir, 'pkl', 'best_' + args.load_experiment + '.npz')
        loaded = dict(numpy.load(path))
        opt = {k: loaded.pop(k) for k in list(loaded) if k.startswith('__adam_')}
        parrot.set_parameter_values(loaded)
        if len(opt) == 3:                                                      # resume the optimizer as well
            algorithm.m.copy_(torch.from_numpy(opt['__adam_m']))
            algorithm.v.copy_(torch.from_numpy(opt['__adam_v']))
            algorithm.time = int(opt['__adam_time'])

    cost_name = args.which_cost
    schedule = None
    if args.lr_schedule:                                                       # train.py:175-181
        adam = step_rule.components[1]
        schedule = LearningRateSchedule(adam, algorithm, parrot,
                                        os.path.join(save_dir, 'pkl', 'best_' + exp_name + '.npz'),
                                        patience=10, num_cuts=5)
    best = float('inf')
    t0 = time.time()
    running, seen = 0.0, 0
    if rank == 0:
        print('Training starting:')
    for it, tup in enumerate(train_stream.get_epoch_iterator()):
        batch = dict(zip(train_stream.sources, tup))
        batch = parallel.shard_batch({('speaker' if k == 'speaker_index' else k): v for k, v in batch.items()},
                                     rank, world)
        if 'speaker' in batch:
            batch['speaker_index'] = batch.pop('speaker')
        algorithm.process_batch(batch, args.batch_size // world)
        running += algorithm.global_cost()
        seen += 1
        done = args.steps is not None and it + 1 >= args.steps
        timed_out = args.time_limit is not None and (time.time() - t0) > args.time_limit * 3600   # TimedFinish
        if world > 1 and args.time_limit is not None:
            # every rank must leave the loop (and enter the validation collectives) at the same iteration
            flag = torch.tensor([1.0 if timed_out else 0.0], device=parrot.device)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MAX)
            timed_out = bool(flag.item() > 0)
        if (it + 1) % args.save_every == 0 or done or timed_out:
            train_cost = running / seen
            running, seen = 0.0, 0
            # validation monitor (train.py:118-123): cost on a few batches, no update, no state carry
            vs = parrot_stream(args.dataset, args.use_speaker, ('valid',), args.batch_size, **valid_stream_args)
            vcost, vn = 0.0, 0
            carried = parrot.get_state()          # the monitor must not disturb the TBPTT state
            for vt in vs.get_epoch_iterator():
                vb = dict(zip(vs.sources, vt))
                vb = parallel.shard_batch({('speaker' if k == 'speaker_index' else k): v for k, v in vb.items()},
                                          rank, world)
                parrot.compute_cost(vb['features'], vb['features_mask'], vb['labels'], vb['labels_mask'],
                                    vb.get('speaker'), vb['start_flag'], args.batch_size // world)
                vcost += algorithm.global_cost()
                vn += 1
                if vn >= 4:
                    break
            vcost /= max(vn, 1)
            parrot.set_state(carried)
            if rank == 0:
                print('iter %d  train_%s %.5f  valid_%s %.5f  (%.1f s)' %
                      (it + 1, cost_name, train_cost, cost_name, vcost, time.time() - t0))
                vals = parrot.get_parameter_values()
                # (Blocks checkpoints carry the algorithm buffers too: Adam moments and step count ride along)
                vals.update({'__adam_m': algorithm.m.cpu().numpy(), '__adam_v': algorithm.v.cpu().numpy(),
                             '__adam_time': numpy.int64(algorithm.time)})
                numpy.savez(os.path.join(save_dir, 'pkl', 'last_' + exp_name + '.npz'), **vals)   # train.py:166-173
                if vcost < best:                                                                  # TrackTheBest
                    best = vcost
                    numpy.savez(os.path.join(save_dir, 'pkl', 'best_' + exp_name + '.npz'), **vals)
            if world > 1:
                torch.distributed.barrier()          # the best-parameter file is complete before anyone reloads it
            if schedule is not None and schedule.do(vcost, rank, world):
                done = True                          # training_finish_requested (extensions.py:148-150)
        if done or timed_out:
            break
    return parrot


if __name__ == '__main__':
    main()
