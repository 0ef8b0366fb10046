# Included from the reference's top-level CMakeLists.txt (integration/patches/0005-cmake-link-the-b200-library.patch) when it
# is configured with -DCRYPTO12381_B200=<root of this repository>.  Defines the `crypto12381` target so that
#   * the reference's own src/miracl_core_interface.cpp is compiled UNMODIFIED and its nine hot definitions are renamed
#     aside with objcopy (rename_hot.py: sum_of_products, double_multiply, multiply on point1 / point2 / fp12, pow, pair_ate,
#     pair_double_ate, pair_final_exponentiation) - the non-hot bridge functions keep forwarding to MIRACL-core;
#   * integration/miracl_core_interface_b200.cpp supplies those nine (and the two ABI-additive entries of SURVEY §8f N2)
#     over libc12381_cuda.so;
#   * every target linking `crypto12381` (examples, unit tests) picks the CUDA library up transitively.
# Expects crypto12381_srcs / miracl_core_srcs as the reference's CMakeLists.txt globs them.  No `-static`: libcudart and
# the driver library are shared objects.
find_package(Python3 COMPONENTS Interpreter REQUIRED)
find_program(C12381_OBJCOPY objcopy REQUIRED)

set(_c12381_root "${CRYPTO12381_B200}")
set(_c12381_lib "${_c12381_root}/crypto12381_b200/libc12381_cuda.so")
if(NOT EXISTS "${_c12381_lib}")
    message(FATAL_ERROR "${_c12381_lib} is missing: build it first (python -m crypto12381_b200.build in ${_c12381_root})")
endif()

list(FILTER crypto12381_srcs EXCLUDE REGEX "miracl_core_interface\\.cpp$")

add_library(crypto12381_stock_bridge OBJECT "${CMAKE_CURRENT_SOURCE_DIR}/src/miracl_core_interface.cpp")
target_include_directories(crypto12381_stock_bridge PRIVATE "${CMAKE_CURRENT_SOURCE_DIR}/include" "${CMAKE_CURRENT_SOURCE_DIR}/3rd-party")

set(_c12381_renamed "${CMAKE_CURRENT_BINARY_DIR}/stock_bridge_renamed.o")
set(_c12381_table "${CMAKE_CURRENT_BINARY_DIR}/stock_bridge_redefine.txt")
add_custom_command(
    OUTPUT "${_c12381_renamed}"
    COMMAND "${Python3_EXECUTABLE}" "${_c12381_root}/integration/rename_hot.py" "$<TARGET_OBJECTS:crypto12381_stock_bridge>" "${_c12381_table}"
    COMMAND "${C12381_OBJCOPY}" "--redefine-syms=${_c12381_table}" "$<TARGET_OBJECTS:crypto12381_stock_bridge>" "${_c12381_renamed}"
    DEPENDS crypto12381_stock_bridge "$<TARGET_OBJECTS:crypto12381_stock_bridge>" "${_c12381_root}/integration/rename_hot.py"
    COMMENT "renaming the stock definitions of the hot bridge functions aside"
    VERBATIM)
set_source_files_properties("${_c12381_renamed}" PROPERTIES EXTERNAL_OBJECT TRUE GENERATED TRUE)

add_library(crypto12381 STATIC ${crypto12381_srcs} ${miracl_core_srcs}
            "${_c12381_root}/integration/miracl_core_interface_b200.cpp" "${_c12381_renamed}")
target_include_directories(crypto12381 PUBLIC "${CMAKE_CURRENT_SOURCE_DIR}/include"
                           PRIVATE "${CMAKE_CURRENT_SOURCE_DIR}/3rd-party" "${_c12381_root}/include")
target_link_libraries(crypto12381 PUBLIC "${_c12381_lib}")
set_target_properties(crypto12381 PROPERTIES INTERFACE_LINK_OPTIONS "-Wl,-rpath,${_c12381_root}/crypto12381_b200")
