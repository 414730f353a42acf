# Stub package: lets `import sksparse.cholmod` inside the reference succeed in an image without CHOLMOD.
