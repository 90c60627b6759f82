"""Worker process of the batched SQP driver (sqp._ProcessStepper): steps SciPy's SLSQP cores of the
instances it owns.  Started as `python -m opengoddard_b200.sqp_worker <socket>`; never touches CUDA.
"""
import os
import sys
from multiprocessing import connection


def main():
    address = sys.argv[1]
    key = bytes.fromhex(os.environ["OGB200_SQP_KEY"])
    from opengoddard_b200 import sqp
    with connection.Client(address, family="AF_UNIX", authkey=key) as conn:
        sqp._worker_loop(conn)


if __name__ == "__main__":
    main()
