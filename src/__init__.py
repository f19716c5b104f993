"""Import-path shim: lets code written against the reference (``from src.models import ...``,
``from src.config import load_config``) run on the B200 implementation unchanged."""
