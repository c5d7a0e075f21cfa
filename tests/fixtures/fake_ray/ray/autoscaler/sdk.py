REQUESTS = []


def request_resources(num_cpus=None, bundles=None):
    REQUESTS.append(list(bundles or []))
