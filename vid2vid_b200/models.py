"""models/models.py:61-101 for one process per GPU: the factory train.py / test.py call.  There is no DataParallel wrapper
(models.py:10-59 scatter the batch over GPUs inside one process; here each rank owns one GPU and `trainer.Trainer` exchanges
gradients with one NCCL all-reduce), so the returned objects are the models themselves -- no `.module` indirection."""


def create_model(opt):
    """models.py:61-84: the generator model for inference; [modelG, modelD, flowNet] for training."""
    if getattr(opt, 'model', 'vid2vid') != 'vid2vid':
        raise ValueError('Model [%s] not recognized.' % opt.model)
    if getattr(opt, 'fp16', False):
        raise NotImplementedError('--fp16 (apex amp) is out of scope: the precise / fast arithmetic modes replace it (DESIGN.md section 4)')
    from .model_g import Vid2VidModelG
    modelG = Vid2VidModelG()
    modelG.initialize(opt)
    if not opt.isTrain:
        return modelG
    from .flownet import FlowNet
    from .model_d import Vid2VidModelD
    modelD = Vid2VidModelD()
    modelD.initialize(opt)
    flowNet = FlowNet()
    flowNet.initialize(opt)
    return [modelG, modelD, flowNet]


def create_optimizer(opt, models):
    """models.py:86-101 (without apex): the optimizers the models built in initialize()."""
    modelG, modelD, flowNet = models
    optimizer_D_T = [getattr(modelD, 'optimizer_D_T' + str(s)) for s in range(opt.n_scales_temporal)]
    return modelG, modelD, flowNet, modelG.optimizer_G, modelD.optimizer_D, optimizer_D_T


def create_trainer(opt, world=1):
    """The training step of train.py:50-93 over create_model's objects (one rank)."""
    from .trainer import Trainer
    modelG, modelD, flowNet = create_model(opt)
    return Trainer(opt, modelG, modelD, flowNet, world=world)
