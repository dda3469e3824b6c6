class ModelV2:
    """ray.rllib.models.modelv2.ModelV2 (1.11): __call__ restores obs into obs_flat and dispatches to forward()."""

    def __init__(self, obs_space, action_space, num_outputs, model_config, name, framework="torch"):
        self.obs_space = obs_space
        self.action_space = action_space
        self.num_outputs = num_outputs
        self.model_config = model_config
        self.name = name or "default_model"
        self.framework = framework
        self._last_output = None

    def __call__(self, input_dict, state=None, seq_lens=None):
        restored = dict(input_dict)
        restored["obs"] = input_dict["obs"]
        restored["obs_flat"] = input_dict["obs"]
        res = self.forward(restored, state or [], seq_lens)
        outputs, state_out = res
        self._last_output = outputs
        return outputs, state_out if len(state_out) > 0 else (state or [])
