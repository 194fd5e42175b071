"""Drop-in import paths of the reference for the single-env API: put this directory on sys.path."""
