"""TEST INFRASTRUCTURE -- NOT idaes-pse: see multiperiod/multiperiod.py"""
__stub__ = True
