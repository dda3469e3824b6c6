import copy


def grid_search(values):
    return {"grid_search": values}


class Trainable:
    """ray.tune.Trainable: the constructor calls setup(config); train() calls step()."""

    def __init__(self, config=None, logger_creator=None):
        self.config = config or {}
        self._iteration = 0
        self.setup(copy.deepcopy(self.config))

    def setup(self, config):
        pass

    def step(self):
        raise NotImplementedError

    def train(self):
        result = self.step()
        self._iteration += 1
        result = dict(result)
        result["training_iteration"] = self._iteration
        return result

    def save_checkpoint(self, checkpoint_dir):
        raise NotImplementedError

    def load_checkpoint(self, checkpoint):
        raise NotImplementedError

    def restore(self, checkpoint_path):
        self.load_checkpoint(checkpoint_path)


def run(*args, **kwargs):
    raise RuntimeError("ray.tune.run is not available under the oracle stub")
