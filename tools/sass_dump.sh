#!/bin/bash
# usage: tools/sass_dump.sh <object-glob-stem e.g. vb_render> <mangled-name regex>  -> prints SASS of the first matching function
obj=$(ls vampire_b200/_lib/obj/$1.*.o | head -1)
fn=$(cuobjdump -elf $obj 2>/dev/null | grep -oE "\.text\.[A-Za-z0-9_]+" | sed 's/.text.//' | sort -u | grep -E "$2" | head -1)
cuobjdump -sass -fun "$fn" $obj | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s#^\s+/\*([0-9a-f]{4})\*/#\1#'
