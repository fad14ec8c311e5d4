"""soket_b200.utils -- mirrors soket/utils (data pipeline)."""
