"""Small manifest fragments."""


def owner_reference_template(namespace, name, uid, kind="AdaptDLJob",
                             api="adaptdl.petuum.com/v1"):
    """ownerReferences entry making ``kind/name`` the controller-owner (pods
    are garbage-collected with their job)."""
    return [{"apiVersion": api, "controller": True,
             "blockOwnerDeletion": True, "kind": kind, "name": name,
             "uid": uid}]
