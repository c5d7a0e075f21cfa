"""``python -m adaptdl_b200.cli ...`` == the ``adaptdl-b200`` console script."""

import sys

from adaptdl_b200.cli.main import main

if __name__ == "__main__":
    sys.exit(main())
