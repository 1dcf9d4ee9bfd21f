"""Mirror of the reference's ``util`` package for the hot-path functions (``util.loss``, ``util.utils``)."""
