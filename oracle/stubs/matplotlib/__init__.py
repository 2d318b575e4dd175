"""Import-only stand-in for matplotlib (TEST INFRASTRUCTURE, used only when the real package is absent).

gpmp/mcmc/mh.py:53 imports matplotlib.pyplot at module level, so `import gpmp.mcmc` needs the name to resolve even
though no test plots anything.  oracle/vendor_ref.py appends this directory to sys.path only when matplotlib is
not installed.  Any attempt to actually draw raises."""
__version__ = "0+stub"
