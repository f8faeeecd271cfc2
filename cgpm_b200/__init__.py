"""cgpm_b200 -- B200-native CrossCat Gibbs inference behind the cgpm Engine/State API."""
__version__ = '0.1.0'
