"""Built-in copies of the reference's six flat configs as Python dicts (same keys and values as
configs/**.yaml of the reference; the YAML files themselves are not shipped).  Any reference YAML
file can still be passed to the entry points - these are for synthetic runs, tests and bench.py."""
import copy

_BASE = dict(
    path2data='./data/', path2save='./runs/', meshes_fname='ShapeNetCore55v2_meshes_resampled.h5', cloud_size=2048,
    chosen_label=0, batch_size=64, shuffle=True, num_workers=8,
    cloud_rescale2orig=False, cloud_recenter2orig=False, cloud_translate=False,
    cloud_translate_shift=[0.00055863, 0.00127477, 0.01701898], cloud_scale=True, cloud_scale_scale=2.0,
    cloud_noise=False, cloud_noise_scale=0.002, cloud_center=False,
    pc_enc_init_n_channels=3, pc_enc_init_n_features=64, pc_enc_n_features=[128, 256, 512],
    deterministic=False, g_latent_space_size=128, g_prior_n_flows=7, g_prior_n_features=128, g_posterior_n_layers=1,
    p_latent_space_size=3, p_prior_n_layers=1, p_decoder_n_flows=21, p_decoder_n_features=64,
    p_decoder_base_type='free', p_decoder_base_var=-3.9551,
    train_mode='p_rnvp_mc_g_rnvp_vae', util_mode='training', pnll_weight=1.0, gnll_weight=1.0, gent_weight=1.0,
    n_epochs=800, resume=False, resume_optimizer=False, min_lr=0.000256, max_lr=0.000256, beta1=0.9,
    min_beta2=0.99, max_beta2=0.99, cycle_length=400, wd=0.000001,
)

_AE = dict(chosen_label=None, g_latent_space_size=512, n_epochs=400, min_beta2=0.995, max_beta2=0.995)

_OVERRIDES = {
    'generation/airplane': {},
    'generation/car': dict(chosen_label=16, p_decoder_base_var=-3.5086, n_epochs=1000, min_beta2=0.95, max_beta2=0.95,
                           cycle_length=500),
    'generation/chair': dict(chosen_label=18, p_decoder_base_var=-3.6990),
    'autoencoding/all_original': dict(_AE, cloud_rescale2orig=True, cloud_recenter2orig=True, cloud_translate=True,
                                      cloud_scale=False, p_decoder_base_var=-3.5960),
    'autoencoding/all_scaled': dict(_AE, p_decoder_base_type='freevar', p_decoder_base_var=-3.5960),
    'svr/all': dict(_AE, p_decoder_base_type='freevar', p_decoder_base_var=0., train_mode='p_rnvp_mc_g_rnvp_vae_ic',
                    images_fname='ShapeNetAll13_images.h5', meshes_fname='ShapeNetAll13_meshes.h5', cloud_size=2500,
                    batch_size=50, g_prior_n_layers=1, n_epochs=20, cycle_length=20,
                    cloud_translate_shift=[-0.01728925, 0.01043933, -0.00012057],
                    image_resize=True, image_size=[224, 224], image_pad=False, image_pad_size=[0, 0],
                    image_add_grayscale=True, image_normalize=True,
                    image_means=[0.03492457, 0.03379815, 0.03475684, 0.03874264, 0.14249628],
                    image_stds=[0.10963749, 0.10795733, 0.11031612, 0.12266339, 0.34554828],
                    image_noise=False, image_noise_scale=0.02, image_remove_alpha=True,
                    img_enc_init_n_channels=4, img_enc_init_n_features=32, img_enc_n_features=[64, 128, 256],
                    img_enc_bottleneck_n_features=256),
}


def names():
    return sorted(_OVERRIDES)


def get(name):
    """Flat config dict, e.g. get('generation/chair')."""
    if name not in _OVERRIDES:
        raise KeyError("unknown config %r (have %s)" % (name, names()))
    cfg = copy.deepcopy(_BASE)
    cfg.update(copy.deepcopy(_OVERRIDES[name]))
    return cfg


def load(path_or_name):
    """A built-in name or a path to a reference-style YAML file (yaml.safe_load: the reference's bare
    yaml.load(stream) fails on PyYAML >= 6)."""
    if path_or_name in _OVERRIDES:
        return get(path_or_name)
    import io
    import yaml
    with io.open(path_or_name, 'r') as f:
        return yaml.safe_load(f)
