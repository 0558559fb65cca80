"""empty stub (environment16.py:13 imports pyplot and never uses it)"""
