from .clip import *  # noqa: F401,F403  (the reference's `import clip` surface)
