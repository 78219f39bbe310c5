"""Mirror of the reference's `common` package (imported there as `import common as lib`)."""
from . import ops  # noqa: F401
