"""Mirror of the reference's `common` package (imported there as `import common as lib`)."""
from . import ops  # noqa: F401


def print_model_settings(locals_):
    """common/__init__.py:56-62"""
    print("Uppercase local vars:")
    all_vars = sorted(((k, v) for (k, v) in locals_.items()
                       if k.isupper() and k not in ('T', 'SETTINGS', 'ALL_SETTINGS')), key=lambda x: x[0])
    for var_name, var_value in all_vars:
        print("\t{}: {}".format(var_name, var_value))


def print_model_settings_dict(settings):
    """common/__init__.py:65-70"""
    print("Settings dict:")
    for var_name, var_value in sorted(settings.items(), key=lambda x: x[0]):
        print("\t{}: {}".format(var_name, var_value))
