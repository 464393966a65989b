"""Mirrors of the Blocks objects train.py wires together (train.py:100-108):

    step_rule = CompositeRule([StepClipping(10. * args.grad_clip), Adam(args.learning_rate)])
    algorithm = GradientDescent(cost=cost, parameters=parameters, step_rule=step_rule)
    algorithm.add_updates(extra_updates)

There is no symbolic graph here, so ``GradientDescent`` drives the device model eagerly:
``process_batch`` = compute_cost -> backward -> (data-parallel allreduce) -> fused clip + Adam
(``parrot_adam_clip_step``, one pass over the flat parameter / gradient / moment buffers).
Blocks >= 0.2 semantics (PARITY UNPINNED, the reference pins no Blocks version): Adam
beta1=0.9, beta2=0.999, epsilon=1e-8, bias-corrected learning rate; StepClipping rescales the
whole step when the global L2 norm exceeds the threshold.
"""
import ctypes as C

import torch

from . import _lib, parallel


class StepClipping(object):
    def __init__(self, threshold=None):
        self.threshold = threshold


class Adam(object):
    def __init__(self, learning_rate=0.002, beta1=0.9, beta2=0.999, epsilon=1e-8, decay_factor=1):
        assert decay_factor == 1, 'decay_factor != 1 is not implemented'
        self.learning_rate = learning_rate
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon


class CompositeRule(object):
    def __init__(self, components):
        self.components = list(components)


class GradientDescent(object):
    """blocks.algorithms.GradientDescent for a ``parrot_b200.Parrot`` model."""

    def __init__(self, cost=None, parameters=None, step_rule=None, model=None, process_group=None,
                 on_unused_sources='warn'):
        assert model is not None, 'pass model=<parrot_b200.Parrot>'
        self.model = model
        rules = step_rule.components if isinstance(step_rule, CompositeRule) else [step_rule]
        self.clip = next((r for r in rules if isinstance(r, StepClipping)), StepClipping(float('inf')))
        self.adam = next((r for r in rules if isinstance(r, Adam)), None)
        assert self.adam is not None, 'only Adam is implemented'
        self.step_rule = step_rule
        self.group = process_group
        model._allocate()
        n = model.num_floats
        dev = model.device
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.stats = torch.zeros(4, dtype=torch.float32, device=dev)
        self.scratch = torch.zeros(1024, dtype=torch.float64, device=dev)
        self.time = 0
        self.extra_updates = []
        self.last_cost = None

    def add_updates(self, updates):
        """train.py:108.  The carried-state updates are applied inside compute_cost already."""
        self.extra_updates = list(updates)

    def zero_buffers(self):
        """extensions.py LearningRateSchedule zeroes the Adam buffers after reloading the best params."""
        self.m.zero_(); self.v.zero_(); self.time = 0

    def step(self):
        """backward of the last compute_cost + allreduce + clip + Adam."""
        model = self.model
        lib = _lib.load()
        model.backward(unnormalised=True)
        flat = parallel.allreduce_flat(model.flat_grads, self.group)
        # the divisor 1/(sum(mask)+1e-5) is read on the device from flat[-1]: no host sync in the step
        self.time += 1
        stream = torch.cuda.current_stream(model.device).cuda_stream
        thr = self.clip.threshold if self.clip.threshold is not None else float('inf')
        _lib.check(lib.parrot_adam_clip_step(
            C.c_void_p(model.flat_params.data_ptr()), C.c_void_p(flat.data_ptr()),
            C.c_void_p(self.m.data_ptr()), C.c_void_p(self.v.data_ptr()), model.num_floats,
            1.0, C.c_void_p(flat.data_ptr() + 4 * model.num_floats), min(thr, 3.0e38),
            self.adam.learning_rate, self.adam.beta1, self.adam.beta2,
            self.adam.epsilon, self.time, C.c_void_p(self.stats.data_ptr()),
            C.c_void_p(self.scratch.data_ptr()), C.c_void_p(stream)))
        model.mark_dirty()

    def process_batch(self, batch, batch_size=None):
        """One training step on a batch dict with the reference's source names
        (features, features_mask, labels, labels_mask[, speaker_index], start_flag)."""
        model = self.model
        B = batch_size or batch['features'].shape[1]
        cost, updates, att, _ = model.compute_cost(
            batch['features'], batch['features_mask'], batch['labels'], batch['labels_mask'],
            batch.get('speaker_index'), batch.get('start_flag', 1.0), B,
            feedback_noise=batch.get('feedback_noise'), noise_level=batch.get('feedback_noise_level'))
        self.step()
        self.last_cost = cost
        return cost

    def global_cost(self):
        """Masked mean over the global batch: allreduced sum(cost*mask) / (sum(mask) + 1e-5)."""
        t = self.model.cost_terms[1:3].clone()
        parallel.allreduce_flat(t, self.group)
        return float(t[0]) / (float(t[1]) + 1e-5)
