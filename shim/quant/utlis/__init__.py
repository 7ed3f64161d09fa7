"""`from quant.utlis import register_activation_hooks` (steps/run_dpd.py:17 imports it; nothing on the step path calls it)."""


def register_activation_hooks(model, layer_types=None):
    """Forward hooks that record each matching sub-module's output; returns (activations dict, hook handles)."""
    acts, handles = {}, []
    for name, mod in model.named_modules():
        if layer_types is None or isinstance(mod, tuple(layer_types)):
            handles.append(mod.register_forward_hook(lambda m, i, o, name=name: acts.__setitem__(name, o)))
    return acts, handles
