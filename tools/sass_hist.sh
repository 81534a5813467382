#!/bin/bash
# Opcode histogram of one kernel of an object file: tools/sass_hist.sh <obj> <mangled-name-regex> [top]
obj="$1"; pat="$2"; top="${3:-30}"
name=$(cuobjdump -sass "$obj" | grep -oE "Function : \S+" | awk '{print $3}' | grep -E "$pat" | head -1)
echo "kernel: $name"
cuobjdump -sass -fun "$name" "$obj" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sort | uniq -c | sort -rn | head -"$top"
cuobjdump -res-usage -fun "$name" "$obj" 2>/dev/null | grep -E "REG|STACK" | head -2
