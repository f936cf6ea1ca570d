#!/usr/bin/env python
"""usage: bamse.py [-h] [-e error] [-nc max_clusts] [-s sparsity] [-n n_trees] [-t top_trees] inputfile [inputfile ...]
Same command line as the reference's bamse.py; the scoring runs on a B200 through libbamse_cuda."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bamse_b200.cli import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
