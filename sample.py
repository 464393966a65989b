"""Sampling entry point (reference sample.py:1-289; model build, sampling call and the
end-of-utterance heuristic sample.py:95-163; plots / animations are out of scope).

    RESULTS_DIR=/tmp/res python sample.py --experiment_name baseline --num_samples 4 --num_steps 200
"""
import os
import pickle

import numpy

from generate import generate_wav
from parrot_b200 import Parrot
from parrot_b200.datasets import SyntheticVoice, parrot_stream
from parrot_b200.utils import sample_parse, stop_heuristic


def main(argv=None):
    args = sample_parse(argv)
    # flags of the reference's parser whose code paths (sample.py:78-93, 120-134: custom phrases through the text
    # front end, speaker mixing, random speakers, the one-step debugging sampler) are not implemented here
    for flag in ('phrase', 'mix', 'random_speaker', 'sample_one_step'):
        if getattr(args, flag, None):
            raise NotImplementedError('sample.py: --%s is parsed for compatibility but not supported' % flag)
    with open(os.path.join(args.save_dir, 'config', args.experiment_name + '.pkl'), 'rb') as f:   # sample.py:24-28
        saved_args = pickle.load(f)
    assert saved_args.dataset == args.dataset
    params_mode = 'last_' if args.use_last else 'best_'
    args.samples_name = params_mode + args.samples_name
    parameters = dict(numpy.load(os.path.join(args.save_dir, 'pkl', params_mode + args.experiment_name + '.npz')))

    dataset = SyntheticVoice(output_dim=saved_args.output_dim, num_characters=saved_args.num_characters,
                             num_speakers=saved_args.num_speakers, seed=args.seed + 7)
    test_stream = parrot_stream(args.dataset, saved_args.use_speaker, ('test',), args.num_samples,
                                args.num_steps, sorting_mult=1, labels_type='text', dataset=dataset)
    data_tr = dict(zip(test_stream.sources, next(iter(test_stream.get_epoch_iterator()))))
    labels_tr, labels_mask_tr = data_tr['labels'], data_tr['labels_mask']
    features_mask_tr = data_tr['features_mask']
    speaker_tr = data_tr.get('speaker_index')
    if args.speaker_id and saved_args.use_speaker:                                    # sample.py:76-77
        speaker_tr = speaker_tr * 0 + args.speaker_id

    parrot_args = {                                                                   # sample.py:95-116
        'input_dim': saved_args.input_dim, 'output_dim': saved_args.output_dim,
        'rnn_h_dim': saved_args.rnn_h_dim, 'readouts_dim': saved_args.readouts_dim,
        'weak_feedback': saved_args.weak_feedback, 'full_feedback': saved_args.full_feedback,
        'feedback_noise_level': None, 'layer_norm': saved_args.layer_norm,
        'use_speaker': saved_args.use_speaker, 'num_speakers': saved_args.num_speakers,
        'speaker_dim': saved_args.speaker_dim, 'which_cost': saved_args.which_cost,
        'num_characters': saved_args.num_characters, 'attention_type': saved_args.attention_type,
        'attention_alignment': saved_args.attention_alignment, 'sampling_bias': args.sampling_bias,
        'sharpening_coeff': args.sharpening_coeff, 'timing_coeff': args.timing_coeff,
        'encoder_type': saved_args.encoder_type, 'raw_output': False, 'name': 'parrot'}
    parrot = Parrot(**parrot_args)
    parrot.initialize()
    parrot.set_parameter_values(parameters)
    print('Successfully loaded the parameters.')

    gen_x, gen_k, gen_w, gen_pi, gen_phi, gen_pi_att = parrot.sample_model(
        labels_tr, labels_mask_tr, features_mask_tr, speaker_tr, args.num_samples, args.num_steps,
        seed=args.seed)                                                               # sample.py:136-138
    print('Successfully sampled the parrot.')
    gen_x = gen_x.swapaxes(0, 1)
    gen_phi = gen_phi.swapaxes(0, 1)
    features_lengths = []
    for idx in range(args.num_samples):                                               # sample.py:147-163
        this_labels_length = int(labels_mask_tr[idx].sum())
        n = stop_heuristic(gen_phi[idx], this_labels_length, args.num_steps)
        if n == args.num_steps:
            print('Its better to increase the number of samples.')
        features_lengths.append(n)

    samples_dir = os.path.join(args.save_dir, 'samples')
    os.makedirs(samples_dir, exist_ok=True)
    norm_info_file = os.path.join(os.environ.get('FUEL_DATA_PATH', ''), args.dataset,
                                  'norm_info_mgc_lf0_vuv_bap_63_MVN.dat')             # sample.py:176-178
    if not os.path.exists(norm_info_file):
        norm_info_file = os.path.join(samples_dir, 'norm_info_identity.dat')
        numpy.stack([numpy.zeros(saved_args.output_dim, numpy.float32),
                     numpy.ones(saved_args.output_dim, numpy.float32)]).tofile(norm_info_file)
        print('norm info file not found; using identity normalisation at', norm_info_file)
    for idx, this_sample in enumerate(gen_x):                                         # sample.py:180-189
        generate_wav(this_sample[:features_lengths[idx]], samples_dir, args.samples_name + '_' + str(idx),
                     sptk_dir=args.sptk_dir, world_dir=args.world_dir, norm_info_file=norm_info_file,
                     do_post_filtering=args.do_post_filtering)
    return gen_x, features_lengths


if __name__ == '__main__':
    main()
