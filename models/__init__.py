"""Drop-in model package for the reference's plugin loader.

``main.get_model(name)`` (reference ``main.py:42-50``) imports ``models.<name.lower()>`` from the
working directory and takes the class ``<name>``; put this repo root on ``sys.path`` (or run
``main.py`` from it) and ``--model_name LightGCN`` / ``LightGCN_Fusion`` resolve to the
B200-native implementations.
"""
