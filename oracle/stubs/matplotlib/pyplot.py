"""See the package docstring: resolves `import matplotlib.pyplot as plt`; every attribute access raises."""


def __getattr__(name):
    raise RuntimeError(f"matplotlib is not installed (oracle/stubs stand-in): pyplot.{name} is unavailable")
