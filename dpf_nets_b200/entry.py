"""Shared implementation of the three entry points (train_ae.py, train_svr.py, evaluate_ae.py) with
the reference's positional arguments (train_ae.py:19-41, evaluate_ae.py:17-44).  Additions:
`--synthetic N` (no ShapeNet/h5py in this image), `--precision`, and multi-GPU when launched with
torchrun (one process per GPU, NCCL; batch sharding for training, row sharding for the sweep)."""
import argparse
import os

import torch
from torch.utils.data import DataLoader

from . import configs as _configs
from . import dist as _dist


def _setup_device():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('dpf_nets_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group('nccl', device_id=dev)
    return dev


def _dataset(config, part, svr, n_synth):
    if n_synth:
        from .lib.datasets.synthetic import SyntheticCloudDataset
        return SyntheticCloudDataset(n_synth, cloud_size=config['cloud_size'], part=part, with_image=svr,
                                     # evaluate_ae.py:49 / train_ae.py: the datasets hand out orig_c / orig_s when either flag asks for them
                                     return_original_scale=bool(config.get('cloud_rescale2orig') or config.get('orig_scale_evaluation')))
    raise SystemExit('real ShapeNet loading is out of scope here (no h5py / data in this image): pass --synthetic N')


def _loader(ds, config, train):
    rank, world = _dist.world()
    sampler = None
    if world > 1 and train:
        sampler = torch.utils.data.distributed.DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=config['shuffle'])
    elif world > 1:
        # evaluation: contiguous un-padded shards (DistributedSampler pads with duplicates, which would enter the
        # all-pairs metrics); evaluating.evaluate gathers the per-rank clouds back in dataset order
        sampler = _dist.ShardSampler(len(ds), rank, world)
    return DataLoader(ds, batch_size=config['batch_size'], shuffle=(train and config['shuffle'] and sampler is None),
                      sampler=sampler, num_workers=0 if getattr(ds, 'n_shapes', None) else config['num_workers'],
                      pin_memory=True, drop_last=train)


def _model(config, svr, dev, precision):
    from .lib.networks.models import Local_Cond_RNVP_MC_Global_RNVP_VAE, Local_Cond_RNVP_MC_Global_RNVP_VAE_IC
    model = (Local_Cond_RNVP_MC_Global_RNVP_VAE_IC if svr else Local_Cond_RNVP_MC_Global_RNVP_VAE)(**config).to(dev)
    model.pc_decoder.precision = precision
    model.pc_encoder.precision = 'fp32' if precision == 'fp32' else 'auto'      # --precision fp32: the exact paths everywhere
    return model


def train_main(svr=False, argv=None):
    from .lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss
    from .lib.networks.optimizers import Adam, LRUpdater, optimizer_state_from_reference
    from .lib.networks.training import train
    from .lib.networks.utils import cnt_params
    ap = argparse.ArgumentParser(description='Model training script. Provide a suitable config.')
    ap.add_argument('config', type=str, help='Path to a YAML config or a built-in name (e.g. generation/chair).')
    ap.add_argument('modelname', type=str)
    ap.add_argument('n_epochs', type=int)
    ap.add_argument('lr', type=float)
    ap.add_argument('--resume', action='store_true')
    ap.add_argument('--resume_optimizer', action='store_true')
    ap.add_argument('--synthetic', type=int, default=0, help='train on N synthetic shapes')
    ap.add_argument('--precision', default='auto')
    ap.add_argument('--batch_size', type=int, default=None)
    ap.add_argument('--path2save', default=None)
    ap.add_argument('--cuda_graph', action='store_true',
                    help='run the training step as two CUDA graphs (forward+loss+backward | optimizer); static batch shapes')
    args = ap.parse_args(argv)
    config = _configs.load(args.config)
    if args.cuda_graph:
        config['cuda_graph'] = True
    config.update(model_name='{0}.pkl'.format(args.modelname), n_epochs=args.n_epochs, min_lr=args.lr, max_lr=args.lr,
                  resume=bool(args.resume), resume_optimizer=bool(args.resume_optimizer))
    if args.batch_size:
        config['batch_size'] = args.batch_size
    if args.path2save:
        config['path2save'] = args.path2save
    dev = _setup_device()
    it = _loader(_dataset(config, 'train', svr, args.synthetic), config, train=True)
    torch.manual_seed(0)
    model = _model(config, svr, dev, args.precision)
    print('Total number of parameters: {}'.format(cnt_params(model.parameters())))
    criterion = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**config).to(dev)
    optimizer = Adam(model.parameters(), lr=config['max_lr'], weight_decay=config['wd'],
                     betas=(config['beta1'], config['max_beta2']), amsgrad=True)
    scheduler = LRUpdater(len(it), **config)
    cur_epoch = cur_iter = 0
    if config['resume']:
        path = os.path.join(config['path2save'], 'models', 'DPFNets', config['model_name'])
        ck = torch.load(path, map_location=dev, weights_only=False)   # protocol-4 pickles need weights_only=False
        cur_epoch, cur_iter = ck['epoch'], ck['iter']
        model.load_state_dict(ck['model_state'])
        if config['resume_optimizer']:
            # stored in the reference's per-tensor layout (also what a reference checkpoint holds)
            optimizer.load_state_dict(optimizer_state_from_reference(model, ck['optimizer_state']))
        print('Model {} loaded.'.format(path))
    for epoch in range(cur_epoch, config['n_epochs']):
        if hasattr(it.sampler, 'set_epoch'):
            it.sampler.set_epoch(epoch)
        train(it, model, criterion, optimizer, scheduler, epoch, cur_iter, **config)
        cur_iter = 0


def evaluate_main(argv=None):
    from .lib.networks.evaluating import evaluate
    from .lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss
    ap = argparse.ArgumentParser(description='Model evaluation script.')
    ap.add_argument('config', type=str)
    ap.add_argument('modelname', type=str)
    ap.add_argument('part', type=str)
    ap.add_argument('cloud_size', type=int)
    ap.add_argument('sampled_cloud_size', type=int)
    ap.add_argument('mode', type=str, help='training | evaluating | generating | predicting')
    ap.add_argument('--orig_scale_evaluation', action='store_true')
    ap.add_argument('--save', action='store_true')
    ap.add_argument('--reps', type=int, default=1)
    ap.add_argument('--synthetic', type=int, default=0)
    ap.add_argument('--precision', default='auto')
    ap.add_argument('--path2save', default=None)
    ap.add_argument('--no_checkpoint', action='store_true', help='evaluate the randomly initialised model')
    args = ap.parse_args(argv)
    config = _configs.load(args.config)
    config.update(model_name='{0}.pkl'.format(args.modelname), part=args.part, cloud_size=args.cloud_size,
                  sampled_cloud_size=args.sampled_cloud_size, util_mode=args.mode,
                  orig_scale_evaluation=bool(args.orig_scale_evaluation), saving=bool(args.save), N_sets=args.reps)
    if args.path2save:
        config['path2save'] = args.path2save
    svr = config['train_mode'] == 'p_rnvp_mc_g_rnvp_vae_ic'
    dev = _setup_device()
    it = _loader(_dataset(config, args.part, svr, args.synthetic), config, train=False)
    model = _model(config, svr, dev, args.precision)
    if not args.no_checkpoint:
        path = os.path.join(config['path2save'], 'models', 'DPFNets', config['model_name'])
        model.load_state_dict(torch.load(path, map_location=dev, weights_only=False)['model_state'])
        print('Model {} loaded.'.format(path))
    criterion = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**config).to(dev)
    rank, world = _dist.world()
    if world > 1:
        # 'generating' draws every cloud from the prior: ranks must not share a noise stream, or each of them
        # would generate the same clouds and the gathered set would hold R copies of everything
        torch.manual_seed(torch.initial_seed() + 7919 * rank)
    return evaluate(it, model, criterion, **config)
