class PlacementGroupSchedulingStrategy(object):
    def __init__(self, placement_group, placement_group_bundle_index=-1,
                 placement_group_capture_child_tasks=None):
        self.placement_group = placement_group
        self.placement_group_bundle_index = placement_group_bundle_index
