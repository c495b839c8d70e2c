"""Shim of python-fire (absent; demo.py:4,64, benchmark.py:5): positional / --flag command line -> function call."""
import sys


def Fire(component, command=None):
    argv = list(sys.argv[1:] if command is None else command)
    args, kwargs = [], {}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a.startswith("--"):
            if "=" in a:
                k, v = a[2:].split("=", 1)
            else:
                k, v = a[2:], argv[i + 1]
                i += 1
            kwargs[k.replace("-", "_")] = v
        else:
            args.append(a)
        i += 1
    return component(*args, **kwargs)
