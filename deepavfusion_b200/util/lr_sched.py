"""LR schedule and parameter groups (mirror of reference util/lr_sched.py:4-24,77-92).  Host-side
arithmetic only; the fused AdamW reads the resulting per-group ``lr`` / ``weight_decay``."""
import math


def adjust_learning_rate(optimizer, epoch, args):
    """lr_sched.py:4-24: linear warm-up + half-cosine, with an extra cosine ramp of the LR multiplier
    for groups flagged ``pretrained`` (pt_warmup_epochs may be the string '300/2', hence eval)."""
    opt = args.opt
    wu = opt.get("warmup_epochs", 0)
    if epoch < wu:
        lr = opt.lr * epoch / wu
    else:
        lr = opt.lr * 0.5 * (1.0 + math.cos(math.pi * (epoch - wu) / (opt.epochs - wu)))
    pt_wu = eval(str(opt.get("pt_warmup_epochs", -1)))
    if epoch < pt_wu:
        pt_scale = (0.5 - 0.5 * math.cos(math.pi * epoch / pt_wu)) * (opt.pt_lr_mult_end - opt.pt_lr_mult_start) + opt.pt_lr_mult_start
    else:
        pt_scale = opt.get("pt_lr_mult_end", 1.0)
    for group in optimizer.param_groups:
        scale = group.get("lr_scale", 1.0)
        group["lr"] = lr * scale * (pt_scale if group.get("pretrained", False) else 1.0)
    return lr


def param_groups_weight_decay(model, weight_decay=1e-5, no_weight_decay_list=()):
    """timm 0.9.2 optim_factory.param_groups_weight_decay: no decay iff ndim <= 1, name ends with
    '.bias', or name listed."""
    skip = set(no_weight_decay_list)
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim <= 1 or name.endswith(".bias") or name in skip) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


def param_groups_pretrained(model, weight_decay=0.05, no_weight_decay_list=(), image_pt=None, audio_pt=None):
    """lr_sched.py:77-92: (no_decay, decay) groups, with the encoder.image / encoder.audio parameters
    split into extra groups flagged ``pretrained`` when those backbones were initialised from a checkpoint."""
    groups = param_groups_weight_decay(model, weight_decay, no_weight_decay_list)
    pt = []
    if image_pt is not None:
        pt += param_groups_weight_decay(model.encoder.image, weight_decay, no_weight_decay_list)
    if audio_pt is not None:
        pt += param_groups_weight_decay(model.encoder.audio, weight_decay, no_weight_decay_list)
    for g in pt:
        g["pretrained"] = True
    taken = {id(p) for g in pt for p in g["params"]}
    for g in groups:
        g["params"] = [p for p in g["params"] if id(p) not in taken]
    return groups + pt


def param_groups_lrd(model, weight_decay=0.05, no_weight_decay_list=(), layer_decay=0.75):
    """lr_sched.py:25-58 (fine-tuning, eval_finetune.py:199-203): one (decay, no_decay) pair of groups per layer id
    of ``model.params_layer_ids()``, each with ``lr_scale = layer_decay ** (top - layer)``; 1-D parameters and listed
    names are not decayed.  Group order = order of first appearance in ``named_parameters()``."""
    layer_of = {id(p): lid for p, lid in model.params_layer_ids()}
    top = max(layer_of.values())
    skip = set(no_weight_decay_list)
    groups = {}
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        lid = layer_of[id(p)]
        plain = p.ndim == 1 or name in skip
        key = (lid, plain)
        if key not in groups:
            groups[key] = {"lr_scale": layer_decay ** (top - lid), "weight_decay": 0.0 if plain else weight_decay, "params": []}
        groups[key]["params"].append(p)
    return list(groups.values())
