import uuid


class Trial(object):
    """``ray.tune.experiment.Trial``: the constructor arguments, status
    constants and checkpoint hook the integration uses."""

    PENDING = "PENDING"
    RUNNING = "RUNNING"
    PAUSED = "PAUSED"
    TERMINATED = "TERMINATED"
    ERROR = "ERROR"

    def __init__(self, trainable_name, config=None, trial_id=None,
                 experiment_tag="", evaluated_params=None,
                 stopping_criterion=None, placement_group_factory=None,
                 **kwargs):
        self.trainable_name = trainable_name
        self.config = dict(config or {})
        self.trial_id = trial_id or uuid.uuid4().hex[:8]
        self.experiment_tag = experiment_tag
        self.evaluated_params = dict(evaluated_params or {})
        self.stopping_criterion = dict(stopping_criterion or {})
        self.placement_group_factory = placement_group_factory
        self.status = Trial.PENDING
        self.last_result = {}
        self.restore_path = None
        self.checkpoint = None

    def on_checkpoint(self, checkpoint):
        self.checkpoint = checkpoint

    def set_status(self, status):
        self.status = status
