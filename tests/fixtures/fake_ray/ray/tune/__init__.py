class PlacementGroupFactory(object):
    """``ray.tune.PlacementGroupFactory``: a list of resource bundles, the
    first one being the trial driver's."""

    def __init__(self, bundles, strategy="PACK", *args, **kwargs):
        self._bundles = [dict(b) for b in bundles]
        self.strategy = strategy

    @property
    def bundles(self):
        return [dict(b) for b in self._bundles]

    @property
    def head_bundle_is_empty(self):
        return not any(v > 0.01 for v in self._bundles[0].values())

    @property
    def required_resources(self):
        total = {}
        for bundle in self._bundles:
            for key, val in bundle.items():
                total[key] = total.get(key, 0) + val
        return total

    def __eq__(self, other):
        return isinstance(other, PlacementGroupFactory) and \
            self._bundles == other._bundles

    def __hash__(self):
        return hash(repr(self._bundles))


class Trainable(object):
    """``ray.tune.Trainable`` as Tune drives it: ``setup`` once,
    ``train()`` -> ``step()``, ``save()`` -> ``save_checkpoint``,
    ``restore(state)`` -> ``load_checkpoint``, ``stop()`` -> ``cleanup``."""

    def __init__(self, config=None, trial_id="trial_0", **kwargs):
        self.config = dict(config or {})
        self.trial_id = trial_id
        self.iteration = 0
        self.setup(self.config)

    def setup(self, config):
        pass

    def train(self):
        result = self.step()
        self.iteration += 1
        result.setdefault("training_iteration", self.iteration)
        return result

    def save(self, checkpoint_dir=None):
        return self.save_checkpoint(checkpoint_dir)

    def restore(self, state):
        self.load_checkpoint(state)

    def stop(self):
        self.cleanup()

    def cleanup(self):
        pass
