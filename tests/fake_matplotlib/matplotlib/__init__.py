"""A STAND-IN for matplotlib, for tests/test_figure_mpl_walker.py only: just enough of the artist API for the walker
(reasoning_image_generation_b200/figure_mpl.py) to run its control flow -- draw order, what it reads from each artist, what it
refuses.  It draws nothing and proves nothing about matplotlib's behaviour; the real pin is tests/golden/make_figure_golden.py."""
__version__ = '0.0-stand-in'
IS_STAND_IN = True
