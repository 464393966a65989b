"""Flag surface of the reference entry points (utils.py:173-327), same names and defaults.

The reference declares boolean flags with ``type=bool``, which argparse turns into "any non-empty
string is True" (``--layer_norm False`` => True).  Here they parse 'false', '0',
'no', '' as False and everything else as True -- the one deliberate deviation, documented.
Additions: ``--synthetic`` (no HDF5 data in this tree), ``--steps`` (bounded runs).
"""
import argparse
import os


def _bool(s):
    return str(s).strip().lower() not in ('', '0', 'false', 'no', 'none')


def _save_dir_default():
    return os.environ.get('RESULTS_DIR', os.path.join(os.getcwd(), 'results'))   # utils.py:246


def train_parse(argv=None):
    """utils.py:173-254."""
    p = argparse.ArgumentParser()
    p.add_argument('--experiment_name', type=str, default='baseline')
    p.add_argument('--encoder_type', type=str, default='bidirectional')
    p.add_argument('--encoder_dim', type=int, default=128)
    p.add_argument('--input_dim', type=int, default=420)
    p.add_argument('--output_dim', type=int, default=63)
    p.add_argument('--rnn_h_dim', type=int, default=1024)
    p.add_argument('--readouts_dim', type=int, default=1024)
    p.add_argument('--weak_feedback', type=_bool, default=False)
    p.add_argument('--full_feedback', type=_bool, default=False)
    p.add_argument('--feedback_noise_level', type=float, default=None)
    p.add_argument('--layer_norm', type=_bool, default=False)
    p.add_argument('--labels_type', type=str, default='full_labels')
    p.add_argument('--which_cost', type=str, default='MSE')
    p.add_argument('--attention_type', type=str, default='graves')
    p.add_argument('--attention_alignment', type=float, default=1.)
    p.add_argument('--num_characters', type=int, default=43)
    p.add_argument('--batch_size', type=int, default=8)
    p.add_argument('--seq_size', type=int, default=50)
    p.add_argument('--save_every', type=int, default=500)
    p.add_argument('--learning_rate', type=float, default=1e-4)
    p.add_argument('--grad_clip', type=float, default=0.9)
    p.add_argument('--lr_schedule', type=_bool, default=False)
    p.add_argument('--load_experiment', type=str, default=None)
    p.add_argument('--raw_output', type=_bool, default=False)
    p.add_argument('--time_limit', type=float, default=None)
    p.add_argument('--use_speaker', type=_bool, default=False)
    p.add_argument('--num_speakers', type=int, default=22)
    p.add_argument('--speaker_dim', type=int, default=128)
    p.add_argument('--dataset', type=str, default='vctk')
    p.add_argument('--save_dir', type=str, default=_save_dir_default())
    # additions
    p.add_argument('--synthetic', type=_bool, default=True, help='use SyntheticVoice (no HDF5 data here)')
    p.add_argument('--steps', type=int, default=None, help='stop after this many batches')
    p.add_argument('--seed', type=int, default=0)
    args = p.parse_args(argv)
    if args.dataset not in args.save_dir:                                   # utils.py:250-251
        args.save_dir = os.path.join(args.save_dir, args.dataset)
    if args.encoder_type in ('none', 'None'):
        args.encoder_type = None
    return args


def sample_parse(argv=None):
    """utils.py:257-327."""
    p = argparse.ArgumentParser()
    p.add_argument('--experiment_name', type=str, default='baseline')
    p.add_argument('--sampling_bias', type=float, default=1.)
    p.add_argument('--timing_coeff', type=float, default=1.)
    p.add_argument('--sharpening_coeff', type=float, default=1.)
    p.add_argument('--num_samples', type=int, default=10)
    p.add_argument('--num_steps', type=int, default=2048)
    p.add_argument('--samples_name', type=str, default='sample')
    p.add_argument('--speaker_id', type=int, default=None)
    p.add_argument('--mix', type=float, default=None)
    p.add_argument('--dataset', type=str, default='vctk')
    p.add_argument('--new_sentences', type=_bool, default=False)
    p.add_argument('--save_dir', type=str, default=_save_dir_default())
    p.add_argument('--sptk_dir', type=str, default=os.environ.get('SPTK_DIR', ''))
    p.add_argument('--world_dir', type=str, default=os.environ.get('WORLD_DIR', ''))
    p.add_argument('--process_originals', type=_bool, default=False)
    p.add_argument('--do_post_filtering', type=_bool, default=False)
    p.add_argument('--animation', type=_bool, default=False)
    p.add_argument('--debug_plot', type=_bool, default=False)
    p.add_argument('--sample_one_step', type=_bool, default=False)
    p.add_argument('--use_last', type=_bool, default=False)
    p.add_argument('--phrase', type=str, default=None)
    p.add_argument('--random_speaker', type=_bool, default=False)
    p.add_argument('--plot_raw', type=_bool, default=False)
    p.add_argument('--seed', type=int, default=0)
    args = p.parse_args(argv)
    if args.dataset not in args.save_dir:
        args.save_dir = os.path.join(args.save_dir, args.dataset)
    return args


def stop_heuristic(phi, labels_length, num_steps):
    """sample.py:147-163: end of utterance = first frame where phi[t, len] beats every
    phi[t, :len-1]; +40 frames of slack, clamped to num_steps."""
    import numpy as np
    try:
        cond = (phi[:, labels_length, None] > phi[:, :labels_length - 1]).all(axis=1)
        t = np.where(cond)[0][0]
        return int(min(num_steps, t + 40))
    except Exception:
        return int(num_steps)
