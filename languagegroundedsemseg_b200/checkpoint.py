"""Import of the reference's checkpoints into the engine's networks (README.md:146-149 publishes
`Res16UNet34C` / `34D` weights trained with MinkowskiEngine).

The engine keeps MinkowskiEngine's state-dict ABI (`kernel [K,Cin,Cout]`, `[Cin,Cout]` for 1x1, `bias [1,Cout]`,
`*.bn.{weight,bias,running_mean,running_var,num_batches_tracked}`), so importing is a matter of the key prefixes the
reference's own loader strips (lib/utils.py:17-45 `load_state_with_same_shape`: 'module.' from DataParallel, 'model.' from
the Lightning wrapper, 'encoder.' from the pre-training wrapper) and of its lenient mode (main.py:109-117: keep what
matches in name AND shape, leave the rest at its initialisation — e.g. a 20-class `final` head loaded into a 200-class
model)."""
import torch

_PREFIXES = ("module.", "model.", "encoder.")


def strip_reference_prefixes(weights):
    """lib/utils.py:20-31: each prefix is tested on the FIRST key and, if present, cut from every key"""
    weights = dict(weights)
    for p in _PREFIXES:
        if weights and next(iter(weights)).startswith(p):
            weights = {k.partition(p)[2]: v for k, v in weights.items()}
    return weights


def match_reference_weights(model, weights):
    """lib/utils.py:17-45: the entries of `weights` whose (prefix-stripped) name exists in `model` with the same shape"""
    state = model.state_dict()
    weights = strip_reference_prefixes(weights)
    return {k: v for k, v in weights.items() if k in state and tuple(v.shape) == tuple(state[k].shape)}


def load_reference_checkpoint(model, checkpoint, lenient=True, map_location="cpu"):
    """Load a reference checkpoint (path, or the already-loaded dict: `{'state_dict': ...}` as written by
    lib/utils.py:48-70 and by Lightning, or a bare state dict) into `model`.

    lenient=True  main.py:109-117 — matching entries are loaded, the rest keep their initialisation;
    lenient=False main.py:119     — `load_state_dict` on the prefix-stripped weights (strict).
    Returns (loaded keys, checkpoint keys that were skipped, model keys left untouched)."""
    state = torch.load(checkpoint, map_location=map_location) if isinstance(checkpoint, (str, bytes)) or hasattr(
        checkpoint, "__fspath__") else checkpoint
    weights = state["state_dict"] if isinstance(state, dict) and "state_dict" in state else state
    stripped = strip_reference_prefixes(weights)
    if lenient:
        matched = match_reference_weights(model, weights)
        merged = model.state_dict()
        merged.update(matched)
        model.load_state_dict(merged)
    else:
        matched = stripped
        model.load_state_dict(stripped)
    from . import minkowski
    minkowski.invalidate_weight_cache()           # derived tensor-core weight operands are stale now
    skipped = sorted(set(stripped) - set(matched))
    untouched = sorted(set(model.state_dict()) - set(matched))
    return sorted(matched), skipped, untouched
