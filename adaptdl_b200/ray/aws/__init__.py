"""Elastic single-job training on a Ray cluster of AWS spot instances
(reference: ``ray/adaptdl_ray/aws``): a controller actor that plays
supervisor + scheduler for one job, worker tasks that run the user's script
with the ADAPTDL_* environment, a spot-termination watcher, and a greedy
replica optimiser."""
