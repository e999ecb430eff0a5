#!/bin/sh

python3 -m robosat_b200.tools "$@"
