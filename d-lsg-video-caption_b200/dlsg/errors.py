class ConfigurationError(Exception):
    """Stand-in for allennlp.common.checks.ConfigurationError (models/allennlp_beamsearch.py:12)."""
