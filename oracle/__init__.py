"""CPU restatement of the reference's detect-orfs scoring path.

TEST INFRASTRUCTURE: the checker for the CUDA path, never the product.
See oracle/oracle_py.py and oracle/rt_oracle.c.
"""
