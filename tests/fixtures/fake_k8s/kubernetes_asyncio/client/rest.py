class ApiException(Exception):
    def __init__(self, status=None, reason=None):
        super().__init__("({}) {}".format(status, reason))
        self.status = status
        self.reason = reason
