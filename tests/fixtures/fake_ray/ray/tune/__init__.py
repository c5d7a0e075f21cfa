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
