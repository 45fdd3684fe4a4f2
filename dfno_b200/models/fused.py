"""Placeholder hook, replaced once the sm_100a engine lands."""


def wants(args, kwargs, backend) -> bool:
    return False
