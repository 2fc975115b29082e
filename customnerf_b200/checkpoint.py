"""Checkpoints in the reference's format (``Trainer_Nerf.save_checkpoint`` / ``load_checkpoint``,
nerf/utils_init_nerf.py:779-901), so that a ``df_epNNNN.pth`` written by the reference loads here and vice versa.

A reference checkpoint is a dict with
    'epoch', 'global_step', 'stats'                      trainer bookkeeping
    'mean_count', 'mean_density'                         occupancy state kept as Python numbers (cuda_ray)
    'model'                                              NeRFNetwork.state_dict(): pos_en.embeddings / pos_en.offsets,
                                                         network.params, density_network.params, rgb_network.params,
                                                         aabb_*, density_grid, density_bitfield, step_counter
    'optimizer'   (full=True)                            torch.optim.Adam(get_params(lr), betas=(0.9, 0.99), eps=1e-15).state_dict()
    'lr_scheduler' (full=True)                           LambdaLR.state_dict(): 'last_epoch' = train steps taken
    'scaler'      (full=True, fp16)                      GradScaler.state_dict(): scale, growth_factor, backoff_factor,
                                                         growth_interval, _growth_tracker
('ema' is never written by the shipped configuration: ema_decay is None.)  The fused train step keeps the optimiser moments
in one flat vector, the LambdaLR epoch and the loss scaler in eight device words (csrc/adam.cuh); this module maps them
to and from the torch objects' state dicts.
"""
import torch


def _scaler_state_dict(fs):
    """torch.cuda.amp.GradScaler.state_dict() of the device-side scaler (growth / backoff factors are the fixed defaults)"""
    if fs.scaler is None:
        return {}
    h = fs.scaler.cpu()
    return {"scale": float(h[0:1].view(torch.float32)[0]), "growth_factor": 2.0, "backoff_factor": 0.5,
            "growth_interval": int(h[6]), "_growth_tracker": int(h[1])}


def _lr_scheduler_state_dict(fs):
    """LambdaLR.state_dict() as torch writes it for the reference's scheduler (main.py:189); the lambda itself is not
    pickled by torch either (lr_lambdas: [None] for a plain function)"""
    epoch = int(fs.scaler[2]) if fs.scaler is not None else int(fs.step_count)
    base = [fs.lr * 10.0, fs.lr, fs.lr, fs.lr]
    sched = fs.sched.cpu()
    decay = float(sched[6]) ** min(epoch / float(sched[7]), 1.0) if float(sched[7]) > 0 else 1.0
    return {"base_lrs": base, "last_epoch": epoch, "verbose": False, "_step_count": epoch + 1,
            "_get_lr_called_within_step": False, "_last_lr": [b * decay for b in base], "lr_lambdas": [None]}


def checkpoint_state(model, fs=None, epoch=0, global_step=0, stats=None, full=False):
    """the dict ``save_checkpoint`` hands to torch.save (:783-812); ``fs``: the FusedTrainStep that owns the optimiser state"""
    if fs is not None:
        fs.flush()
    state = {"epoch": int(epoch), "global_step": int(global_step),
             "stats": stats if stats is not None else {"loss": [], "valid_loss": [], "results": [], "checkpoints": [],
                                                       "best_result": None}}
    if model.cuda_ray:
        state["mean_count"] = model.mean_count
        state["mean_density"] = model.mean_density
    if full and fs is not None:
        state["optimizer"] = fs.optimizer_state_dict()
        state["lr_scheduler"] = _lr_scheduler_state_dict(fs)
        if fs.scaler is not None:
            state["scaler"] = _scaler_state_dict(fs)
    state["model"] = model.state_dict()
    return state


def save_checkpoint(path, model, fs=None, epoch=0, global_step=0, stats=None, full=False):
    torch.save(checkpoint_state(model, fs, epoch, global_step, stats, full), path)


def load_checkpoint(path_or_state, model, fs=None, model_only=False, log=None):
    """``Trainer_Nerf.load_checkpoint`` (:833-901): a bare state dict or a full checkpoint; strict=False for the model; the
    occupancy numbers; unless ``model_only`` also optimiser moments / step, LambdaLR epoch and loss-scaler state into ``fs``.
    Returns {'epoch', 'global_step', 'stats', 'missing_keys', 'unexpected_keys'}."""
    log = log or (lambda msg: None)
    ckpt = path_or_state
    if not isinstance(ckpt, dict):
        ckpt = torch.load(path_or_state, map_location=next(model.parameters()).device, weights_only=False)
    out = {"epoch": 0, "global_step": 0, "stats": None, "missing_keys": [], "unexpected_keys": []}
    if fs is not None:
        fs.flush()
    if "model" not in ckpt:
        _load_model_state(model, ckpt, strict=True)
        log("[INFO] loaded model.")
        if fs is not None:
            fs._pack()
        return out
    missing, unexpected = _load_model_state(model, ckpt["model"], strict=False)
    out["missing_keys"], out["unexpected_keys"] = list(missing), list(unexpected)
    log("[INFO] loaded model.")
    if missing:
        log("[WARN] missing keys: %s" % (missing,))
    if unexpected:
        log("[WARN] unexpected keys: %s" % (unexpected,))
    if model.cuda_ray:
        if "mean_count" in ckpt:
            model.mean_count = ckpt["mean_count"]
        if "mean_density" in ckpt:
            model.mean_density = ckpt["mean_density"]
    if fs is not None:
        fs._pack()                          # the fp16 operand images of the MLP weights follow the loaded parameters
    if model_only:
        return out
    out["stats"], out["epoch"], out["global_step"] = ckpt.get("stats"), ckpt.get("epoch", 0), ckpt.get("global_step", 0)
    log("[INFO] load at epoch %s, global step %s" % (out["epoch"], out["global_step"]))
    if fs is None:
        return out
    if "optimizer" in ckpt:
        try:
            fs.load_optimizer_state_dict(ckpt["optimizer"])
            log("[INFO] loaded optimizer.")
        except Exception as e:              # the reference swallows a mismatching optimiser state too (:882-886)
            log("[WARN] Failed to load optimizer. (%s)" % (e,))
    if "lr_scheduler" in ckpt and fs.scaler is not None:
        fs.scaler[2] = int(ckpt["lr_scheduler"].get("last_epoch", 0))       # LambdaLR epoch = the scaler's iteration count
        fs.scaler[4:6].zero_()
        log("[INFO] loaded scheduler.")
    if "scaler" in ckpt and ckpt["scaler"] and fs.scaler is not None:
        sd = ckpt["scaler"]
        if float(sd.get("growth_factor", 2.0)) != 2.0 or float(sd.get("backoff_factor", 0.5)) != 0.5:
            log("[WARN] the device-side scaler uses growth 2.0 / backoff 0.5; the checkpoint's factors are ignored")
        fs.scaler[0:1].view(torch.float32).fill_(float(sd["scale"]))
        fs.scaler[1] = int(sd.get("_growth_tracker", 0))
        fs.scaler[6] = int(sd.get("growth_interval", 2000))
        log("[INFO] loaded scaler.")
    return out


def _load_model_state(model, sd, strict):
    """load_state_dict that keeps the parameters where they are (FusedTrainStep's flat vector / peer memory: the Parameters
    are views into it, and load_state_dict copies in place) and accepts a bit field / step counter saved on another device"""
    res = model.load_state_dict(sd, strict=strict)
    return (res.missing_keys, res.unexpected_keys)
