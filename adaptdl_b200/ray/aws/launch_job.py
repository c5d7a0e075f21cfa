"""``adaptdl-b200-on-ray-aws``: run a training script as an elastic job on
a Ray cluster (reference: ``aws/launch_job.py``, console script
``adaptdl_on_ray_aws``).

    python -m adaptdl_b200.ray.aws.launch_job -f train.py -m 8 --gpus 1 \
        --cluster-address auto -- --epochs 10
"""

import argparse
import os


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-f", "--file", required=True)
    parser.add_argument("-m", "--max-cluster-size", type=int, default=None)
    parser.add_argument("-u", "--uri", "--cluster-address", default="auto")
    parser.add_argument("-d", "--working-dir", default=".")
    parser.add_argument("--cpus", type=int, default=1)
    parser.add_argument("--gpus", type=int, default=0)
    parser.add_argument("--port-offset", type=int, default=0)
    parser.add_argument("--checkpoint-timeout", type=int, default=120)
    parser.add_argument("--rescale-timeout", "--cluster-rescale-timeout",
                        dest="rescale_timeout", type=int, default=60)
    parser.add_argument("arguments", nargs=argparse.REMAINDER)
    args = parser.parse_args(argv)
    script_args = args.arguments[1:] if args.arguments[:1] == ["--"] \
        else args.arguments
    if not os.path.exists(os.path.join(args.working_dir, args.file)):
        raise SystemExit("{} not found under {}".format(args.file,
                                                        args.working_dir))
    from adaptdl_b200.ray import require_ray
    ray = require_ray()
    from adaptdl_b200.ray.aws.controller import make_controller_actor
    ray.init(address=args.uri, runtime_env={"working_dir": args.working_dir})
    Controller = make_controller_actor()
    size = args.max_cluster_size or len(ray.nodes())
    controller = Controller.options(name="AdaptDLController").remote(
        size, args.rescale_timeout)
    controller.run_controller.remote()
    resources = {"CPU": args.cpus}
    if args.gpus:
        resources["GPU"] = args.gpus
    status = ray.get(controller.create_job.remote(
        worker_resources=resources, worker_port_offset=args.port_offset,
        checkpoint_timeout=args.checkpoint_timeout, path=args.file,
        argv=script_args))
    print("job finished with status", status)
    return 0 if status == 1 else 1


if __name__ == "__main__":
    raise SystemExit(main())
