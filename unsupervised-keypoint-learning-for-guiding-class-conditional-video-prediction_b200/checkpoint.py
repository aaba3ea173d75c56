"""Checkpoint interchange: name-keyed `.npz` files with the TF variable names of the reference graph.

The reference saves `tf.train.Saver(tf.global_variables())` checkpoints (models/base_model.py:74-81) and restores by NAME,
taking only the variables present in both the file and the graph (:83-92).  TensorFlow cannot run here, so the binary
TF checkpoint format is out of reach; what interchanges is the variable dictionary itself: every array is stored under the
name TF gives it in the reference graph,

    <scope>/conv2d/{kernel,bias}, <scope>/{gamma,beta,moving_mean,moving_variance}      (kernels HWIO)
    <variable>/Adam, <variable>/Adam_1                                                  Adam slots m, v
    beta1_power, beta2_power          non-slot variables of the FIRST optimizer  (train_op_D, detector_translator_model.py:198)
    beta1_power_1, beta2_power_1      ... of the second one                      (train_op_G, :200-202)
    global_step                                                                   (train.py:30)

so a `tf.train.NewCheckpointReader` dump of a real stage-1 checkpoint (`{n: reader.get_tensor(n) for n in ...}` ->
`np.savez`) loads here unchanged, and `export`ed files can be assigned back in TF with one `tf.assign` per name.
beta*_power hold beta^(t+1) after t optimizer steps (TF initialises them to beta and multiplies after every apply).
"""
import math

import numpy as np
import torch

BETA1, BETA2 = 0.5, 0.999


def export_variables(model):
    """All variables of a model (and, for the trainer, the optimizer state) as {tf_name: float32/int64 ndarray}."""
    ctx = model.ctx
    out = {}
    for grp in (ctx.G, ctx.D, ctx.S, ctx.V):
        if grp.data is None:
            continue
        for n in grp.names():
            out[n] = grp.p(n).detach().float().cpu().numpy()
    if hasattr(model, "t_D"):
        for grp in (ctx.D, ctx.G):
            for n in grp.names():
                out[n + "/Adam"] = grp._view(grp.m, n).detach().float().cpu().numpy()
                out[n + "/Adam_1"] = grp._view(grp.v, n).detach().float().cpu().numpy()
        out["beta1_power"] = np.float32(BETA1 ** (model.t_D + 1))
        out["beta2_power"] = np.float32(BETA2 ** (model.t_D + 1))
        out["beta1_power_1"] = np.float32(BETA1 ** (model.t_G + 1))
        out["beta2_power_1"] = np.float32(BETA2 ** (model.t_G + 1))
        out["global_step"] = np.int64(int(model.global_step.value))
        # exact step counters (beta2^t loses t in float32 after a few thousand steps); ignored by TF-side consumers
        out["kp_b200/t_D"] = np.int64(model.t_D)
        out["kp_b200/t_G"] = np.int64(model.t_G)
    return out


def save_npz(path, variables):
    with open(path, "wb") as fh:
        np.savez(fh, **variables)
    return path


def load_npz(path):
    """Plain arrays only: allow_pickle stays False (a checkpoint is data, never code)."""
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def _steps_from_power(power, beta):
    power = float(power)
    if not (0.0 < power < 1.0):
        return 0
    return max(0, int(round(math.log(power) / math.log(beta))) - 1)


def import_variables(model, variables):
    """Assign every array whose name the model knows (others are ignored, like BaseModel.restore of the reference).
    Returns the list of names loaded."""
    ctx = model.ctx
    loaded = list(ctx.load_state_dict({k: v for k, v in variables.items() if ctx.has(k)}))
    if hasattr(model, "t_D"):
        for grp in (ctx.D, ctx.G):
            for n in grp.names():
                for suffix, buf in (("/Adam", grp.m), ("/Adam_1", grp.v)):
                    if n + suffix in variables:
                        grp._view(buf, n).copy_(torch.as_tensor(np.asarray(variables[n + suffix])).to(ctx.device, torch.float32))
                        loaded.append(n + suffix)
        if "kp_b200/t_D" in variables:
            model.t_D, model.t_G = int(variables["kp_b200/t_D"]), int(variables["kp_b200/t_G"])
        else:
            if "beta1_power" in variables:
                model.t_D = _steps_from_power(variables["beta1_power"], BETA1)
            if "beta1_power_1" in variables:
                model.t_G = _steps_from_power(variables["beta1_power_1"], BETA1)
        for n in ("beta1_power", "beta2_power", "beta1_power_1", "beta2_power_1"):
            if n in variables:
                loaded.append(n)
        if "global_step" in variables:
            model.global_step.value = int(variables["global_step"])
            loaded.append("global_step")
    return loaded
