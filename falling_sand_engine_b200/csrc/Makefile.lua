# Builds ../libfse_lua.so = the reference's vendored Lua 5.4.4 (compiled from /root/reference/source/libs/lua WHERE IT LIES, nothing is
# copied into this repository) + the front-door shim lua_front.c.  The .so is git-ignored like every built artefact and travels to the
# GPU box with the snapshot; boxes without /root/reference use the prebuilt file.
LUA_SRC ?= /root/reference/source/libs/lua
CC ?= gcc
CFLAGS ?= -O2 -fPIC -fvisibility=hidden -DLUA_USE_POSIX -DLUA_COMPAT_5_3
OUT := ../libfse_lua.so
LUA_CORE := lapi lcode lctype ldebug ldo ldump lfunc lgc llex lmem lobject lopcodes lparser lstate lstring ltable ltm lundump lvm lzio \
            lauxlib lbaselib lcorolib ldblib liolib lmathlib loadlib loslib lstrlib ltablib lutf8lib linit
OBJS := $(patsubst %,_build/lua_%.o,$(LUA_CORE)) _build/lua_front.o

all: $(OUT)

_build/lua_%.o: $(LUA_SRC)/%.c
	@mkdir -p _build
	$(CC) $(CFLAGS) -I$(LUA_SRC) -c $< -o $@

_build/lua_front.o: lua_front.c
	@mkdir -p _build
	$(CC) $(CFLAGS) -Wall -I$(LUA_SRC) -c $< -o $@

$(OUT): $(OBJS)
	$(CC) -shared -o $@ $(OBJS) -lm -ldl
