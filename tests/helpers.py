"""Parity tests share their workload set-up with bench.py / tools: the code lives in tools/workload_setup.py."""
from tools.workload_setup import *  # noqa: F401,F403
