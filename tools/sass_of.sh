#!/bin/bash
# usage: tools/sass_of.sh <object-or-so> <kernel-name-substring>   -> SASS (address + instruction) of matching kernels
OBJ=$1; PAT=$2
cuobjdump -sass "$OBJ" | awk -v pat="$PAT" '
/Function :/ { on = (index($0, pat) > 0); if (on) print $0 }
on && /^[ \t]+\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\// { line=$0; sub(/^[ \t]+\/\*/,"",line); sub(/\*\/[ \t]+/," ",line); sub(/[ \t]*\/\*.*$/,"",line); print line }'
