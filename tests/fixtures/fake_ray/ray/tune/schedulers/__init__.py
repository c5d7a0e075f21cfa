class TrialScheduler(object):
    """``ray.tune.schedulers.TrialScheduler``: the decision constants and
    the callback names the trial runner invokes."""

    CONTINUE = "CONTINUE"
    PAUSE = "PAUSE"
    STOP = "STOP"
    NOOP = "NOOP"

    def on_trial_add(self, trial_runner, trial):
        raise NotImplementedError

    def on_trial_error(self, trial_runner, trial):
        raise NotImplementedError

    def on_trial_result(self, trial_runner, trial, result):
        raise NotImplementedError

    def on_trial_complete(self, trial_runner, trial, result):
        raise NotImplementedError

    def on_trial_remove(self, trial_runner, trial):
        raise NotImplementedError

    def choose_trial_to_run(self, trial_runner):
        raise NotImplementedError

    def debug_string(self):
        raise NotImplementedError
